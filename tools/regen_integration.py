"""Rewrites the struct declarations embedded in INTEGRATION.md section 1 from b200sr._lib (run after any struct change;
tests/test_abi_cpu.py fails while the document is stale)."""
import os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "remote-sensing-vision-language-diffusion-model_b200"))
from b200sr import _lib
p = os.path.join(ROOT, "INTEGRATION.md")
s = open(p).read()
a = s.index("class Epilogue(C.Structure):")
b = s.index("lib.b200sr_gemm_bf16.restype")
s = s[:a] + _lib.binding_snippet() + "\n" + s[b:]
s = re.sub(r"`b200sr_abi_version\(\)` \(\d+\)", f"`b200sr_abi_version()` ({_lib.ABI_VERSION})", s)
open(p, "w").write(s)
print("INTEGRATION.md regenerated")
