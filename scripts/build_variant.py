"""Build a tuning variant of the library: python scripts/build_variant.py <out.so> [-DMACRO=1 ...] (then run anything
with XVR_B200_LIB=<out.so>)."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ.pop("XVR_B200_LIB", None)
os.environ["XVR_B200_NVCC_FLAGS"] = " ".join(sys.argv[2:])
from xvr_b200 import _build  # noqa: E402

_build.LIB = Path(sys.argv[1]).resolve()
print(_build.build(force=True))
