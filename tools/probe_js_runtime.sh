#!/usr/bin/env bash
# Is there ANY JavaScript/TypeScript/WASM runtime on this machine that could execute the reference
# (/root/reference/src/*.ts) and so generate reference-run golden fixtures?  
# Run here and on the GPU box (gpurun); the output is committed under profiles/.
echo "# JS / TS / WASM runtime probe on $(hostname) at $(date -u +%FT%TZ)"
echo "# image: $(cat /etc/os-release 2>/dev/null | grep PRETTY_NAME | cut -d= -f2)"
for exe in node nodejs bun deno tsc ts-node tsx npx npm yarn pnpm qjs quickjs d8 js js102 js115 gjs rhino jjs graaljs wasmtime wasmer wasm3 wasm-interp cargo rustc emcc; do
  p=$(command -v "$exe" 2>/dev/null)
  if [ -n "$p" ]; then echo "$exe: $p ($("$exe" --version 2>&1 | head -1))"; else echo "$exe: absent"; fi
done
echo "# files named node / bun / deno / *.wasm runtimes anywhere on the filesystem (excluding /proc, /sys, the repo):"
find / -xdev \( -path /proc -o -path /sys -o -path /root/repo -o -path /root/reference \) -prune -o -type f \
  \( -name node -o -name nodejs -o -name bun -o -name deno -o -name qjs -o -name d8 -o -name wasmtime -o -name wasmer -o -name 'libnode.so*' -o -name 'libv8*.so*' -o -name 'libmozjs*.so*' -o -name 'libjavascriptcoregtk*.so*' \) -print 2>/dev/null | head -20
echo "# python packages that embed a JS or WASM engine:"
python - <<'PY'
import importlib.util
for m in ("py_mini_racer", "mini_racer", "quickjs", "js2py", "dukpy", "execjs", "wasmtime", "wasmer", "pywasm", "wasm3", "playwright", "selenium", "pyppeteer", "nodejs", "nodejs_wheel"):
    print(f"{m}: {'PRESENT' if importlib.util.find_spec(m) else 'absent'}")
PY
echo "# /root/reference present: $([ -d /root/reference ] && echo yes || echo no)"
