"""Builds experimental variants of libpyatm_b200.so (different -D flags) into pyatmosphere_b200/variants/ (git-ignored, shipped
to the GPU box by gpurun), for tools/gpu/fft_variants.py.

    python tools/gpu/build_variants.py name1="-DFOO=1 -DBAR" name2="..."
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    env = dict(os.environ, PYATM_LIB=os.path.join(ROOT, "pyatmosphere_b200", "variants", f"libpyatm_{name}.so"),
               PYATM_OBJ_DIR=f"_build_{name}", PYATM_NVCC_FLAGS=flags)
    r = subprocess.run([sys.executable, "-m", "pyatmosphere_b200.build"], cwd=ROOT, env=env, capture_output=True, text=True)
    print(name, flags, "->", r.stdout.strip()[-80:], r.stderr.strip()[-400:])
    if r.returncode:
        sys.exit(1)
