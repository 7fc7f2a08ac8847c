"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/bbq_b200.h declares.
No compute calls (no GPU here); on a box without a GPU bbq_create must fail loudly, never fall back."""
import ctypes as C
import os
import re
import subprocess

import pytest

import bbq_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "bbq_b200.h"), encoding="utf-8").read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(bbq_[a-z0-9_]+)\s*\(", hdr)))


def test_library_builds_and_exports_every_declared_symbol():
    path = bbq_b200.build_library()
    assert os.path.exists(path)
    names = _declared()
    assert len(names) >= 20
    L = C.CDLL(path)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/bbq_b200.h but not exported"
    # and the Python binding table covers exactly the header
    assert sorted(bbq_b200._native.SYMBOLS) == names
    assert bbq_b200._native.load().bbq_abi_version() == 3


def test_library_holds_sm100a_code():
    out = subprocess.run(["cuobjdump", "-lelf", bbq_b200._native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_config_validation_precedes_device_probe():
    # src/binaryQuantizationFormat.ts:143-148
    with pytest.raises(bbq_b200.BbqError) as e:
        bbq_b200.createBinaryQuantizationFormat({"queryBits": 9, "quantizer": {"similarityFunction": "COSINE"}})
    assert e.value.status == 1 and str(e.value) == "queryBits必须在1-8之间"
    with pytest.raises(bbq_b200.BbqError) as e:
        bbq_b200.createBinaryQuantizationFormat({"indexBits": 0, "quantizer": {"similarityFunction": "COSINE"}})
    assert e.value.status == 2 and str(e.value) == "indexBits必须在1-8之间"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(bbq_b200.BbqError) as e:
        bbq_b200.createBinaryQuantizationFormat()
    assert e.value.status == 100


def test_product_never_imports_the_oracle():
    """Nothing under the package may import, include, link or dlopen anything under oracle/."""
    pat = re.compile(r"(from\s+oracle|import\s+oracle|oracle\.oracle|libbbq_oracle|[\"'<(/]oracle/|bbq_oracle)")
    pkg = os.path.join(ROOT, "better-binary-quantization_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cc", ".ts", ".js", ".json")):
                src = open(os.path.join(dp, f), encoding="utf-8").read()
                assert not pat.search(src), f"{os.path.join(dp, f)} references the oracle"


def test_ctypes_structs_mirror_the_header(tmp_path):
    """The ctypes mirrors of bbq_config / bbq_stats (what every Python caller and the tests pass across the C ABI) have
    the header's size and field offsets: a tiny C program compiled against include/bbq_b200.h prints them."""
    fields = {"bbq_config": [f for f, _ in bbq_b200._native.BbqConfig._fields_],
              "bbq_stats": [f for f, _ in bbq_b200._native.BbqStats._fields_]}
    cname = {"lambda_": "lambda"}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "bbq_b200.h"', "int main(void) {"]
    for st, fs in fields.items():
        lines.append(f'  printf("{st} %zu\\n", sizeof({st}));')
        for f in fs:
            lines.append(f'  printf("{st}.{f} %zu\\n", offsetof({st}, {cname.get(f, f)}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for st, cls in (("bbq_config", bbq_b200._native.BbqConfig), ("bbq_stats", bbq_b200._native.BbqStats)):
        assert int(got[st]) == C.sizeof(cls), st
        for f, _ in cls._fields_:
            assert int(got[f"{st}.{f}"]) == getattr(cls, f).offset, (st, f)
