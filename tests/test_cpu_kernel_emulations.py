"""CPU transliterations of kernel logic (several were written before the kernel had run on a GPU):
each script replays a kernel's control flow / index arithmetic thread by thread and compares it with a brute force or
with the oracle's bit-exact reconstruction.  They are evidence about the *algorithms*; the CUDA sources are gated by
the GPU tests (tests/test_*_gpu.py)."""

import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, *args):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script), *args], capture_output=True, text=True,
                       timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_staged_brick_kernel_consumes_every_sample_once_and_reads_the_right_voxels():
    out = _run("emulate_staged_kernel.py")  # asserts internally; prints one line per scene
    served = [float(x) for x in re.findall(r"served from the buffer: ([0-9.]+)", out)]
    assert len(served) == 2 and served[0] == 1.0 and served[1] > 0.85


def test_siddon_integer_walk_reproduces_the_reference_indices():
    out = _run("emulate_siddon_walk.py")
    m = re.search(r"segments with length (\d+) served by the walk ([0-9.]+) wrong (\d+)", out)
    assert m and int(m.group(1)) > 40_000 and float(m.group(2)) > 0.98 and int(m.group(3)) == 0


def test_gather_kernel_window_logic_is_exact_with_the_source_inside_the_volume():
    out = _run("emulate_gather_kernel.py", "3")
    m = re.search(r"sum \|err\| ([0-9.]+) bad voxels (\d+)", out)
    assert m and float(m.group(1)) < 1e-6 and int(m.group(2)) == 0


def test_empty_space_trimming_never_drops_a_sample_that_touches_density():
    """The brick distance field (cell flags -> grown bricks -> separable Chebyshev transform, checked against its
    definition) and the two-ended sphere-tracing walk, emulated in fp32 over 18 000 random rays (sources far, near and
    inside; axis-parallel and plane-grazing rays) through random sparse volumes with odd shapes: every sample with a
    non-zero corner lies inside the trimmed range."""
    out = _run("emulate_trim_walk.py")
    m = re.search(r"rays (\d+) missed samples (\d+)", out)
    assert m and int(m.group(1)) >= 18_000 and int(m.group(2)) == 0
