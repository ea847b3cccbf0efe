"""Probe build of the library next to the product build: python tools/build_probe.py [name] [-Dflag ...]
default: build_variants/libtcprobe.so with -DGM_TC_PROBES=1 (timing probes of gemm_sm100.cu, GM_TC_DEBUG / tools/tc_trace.py).
Select it at run time with GM_LIB_PATH=build_variants/<name>.so."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "libtcprobe"
flags = [a for a in sys.argv[1:] if a.startswith("-")] or ["-DGM_TC_PROBES=1"]
os.environ["GM_NVCC_EXTRA"] = " ".join(flags)
from graph_marl_b200 import build as b  # noqa: E402
out = os.path.join(ROOT, "build_variants")
os.makedirs(out, exist_ok=True)
b.LIBDIR = os.path.join(out, name + "_lib")
b.OBJDIR = os.path.join(out, name + "_obj")
b.LIB = os.path.join(out, name + ".so")
print(b.build(force=True))
