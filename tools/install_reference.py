"""Copy the reference's own Python for the hot path into baseline/_ref (git-ignored; it travels to the GPU box with the repo
snapshot) so that bench.py can time the reference's unmodified rl/algos/ppo.py sample_parallel on the bench host's cores.

    python tools/install_reference.py            (build container only: /root/reference does not exist on the GPU box)

The reference is a script tree without setup.py / pyproject.toml, so there is nothing to pip-install; what is copied is the
source the CPU path imports (rl/, cassie/ without its closed libcassiemujoco.so and the visualiser assets, util/).  Nothing
under baseline/_ref is imported by apex_b200 or committed to this repo.
"""
import os
import shutil
import sys

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")


def main():
    if not os.path.isdir(REF):
        print("no /root/reference here: nothing installed")
        return 1
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST)
    skip = shutil.ignore_patterns("__pycache__", "*.so", "*.stl", "*.png", "*.mp4", "*.gif", "trained_models", ".git")
    for name in ("rl", "cassie", "util"):
        shutil.copytree(os.path.join(REF, name), os.path.join(DST, name), ignore=skip)
    size = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(DST) for f in fs)
    print(f"installed rl/, cassie/, util/ into {DST} ({size / 1e6:.1f} MB)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
