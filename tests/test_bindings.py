"""The reference-side binding: the N-API shim must compile against the C ABI header (syntax check
with the minimal N-API declarations — there is no Node in the image), call only functions include/bbq_b200.h declares,
and the TypeScript class must call only functions the shim registers."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAPI = os.path.join(ROOT, "better-binary-quantization_b200", "bindings", "napi")
TS = os.path.join(ROOT, "better-binary-quantization_b200", "bindings", "ts")


def test_napi_shim_compiles_against_the_header():
    out = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", NAPI, "-I", os.path.join(ROOT, "include"),
                          os.path.join(NAPI, "bbq_napi.c")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_shim_calls_only_declared_entry_points():
    header = open(os.path.join(ROOT, "include", "bbq_b200.h")).read()
    declared = set(re.findall(r"\b(bbq_[a-z_0-9]+)\s*\(", header))
    shim = open(os.path.join(NAPI, "bbq_napi.c")).read()
    called = set(re.findall(r"\b(bbq_[a-z_0-9]+)\s*\(", shim))
    assert called and called <= declared, called - declared
    # the path's entry points are all reachable from JavaScript
    for fn in ("bbq_create", "bbq_index_build", "bbq_search", "bbq_index_export", "bbq_index_attach_rows",
               "bbq_search_rerank", "bbq_index_save", "bbq_index_load", "bbq_destroy", "bbq_index_destroy"):
        assert fn in called, fn


def test_typescript_class_uses_registered_addon_functions():
    shim = open(os.path.join(NAPI, "bbq_napi.c")).read()
    registered = set(re.findall(r'\{"([A-Za-z]+)", NULL, n_', shim))
    ts = open(os.path.join(TS, "binaryQuantizationFormat.gpu.ts")).read()
    used = set(re.findall(r"\baddon\.([A-Za-z]+)\(", ts))
    assert used and used <= registered, used - registered
    # the reference's public methods are all there (src/binaryQuantizationFormat.ts:165,308,583; src/types.ts:32-49)
    for name in ("quantizeVectors", "searchNearestNeighbors", "getConfig", "dimension()", "size()", "getCentroid"):
        assert name in ts, name


def test_error_table_covers_every_reference_status():
    errors = open(os.path.join(TS, "errors.ts")).read()
    header = open(os.path.join(ROOT, "include", "bbq_b200.h")).read()
    for status in range(1, 9):   # the statuses that stand for a `throw new Error(...)` of the reference
        assert re.search(rf"BBQ_ERR_[A-Z_]+ = {status},", header), status
        assert f"case {status}:" in errors, status
