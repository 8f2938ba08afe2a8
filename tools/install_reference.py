"""Copy the UNMODIFIED reference (pure Python, no build step) into the git-ignored baseline/_ref/ so that it travels to
the GPU box with the gpurun snapshot:

    baseline/_ref/models, configs, data      the reference packages (bench.py --impl reference, cpu_baseline)
    baseline/_ref/reftests/test_*.py         the reference's own three acceptance test files, kept apart from the
                                             packages so that `models` resolves to whatever is first on sys.path
    baseline/_ref/MANIFEST.json              sha256 of every copied file

Run in the build container (needs /root/reference):  python tools/install_reference.py
`pip install` is not applicable: the reference has no setup.py / pyproject.toml (requirements.txt only).
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("ERV_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def main():
    if not os.path.isdir(os.path.join(SRC, "models")):
        print(f"{SRC} not found: nothing installed", file=sys.stderr)
        return 1
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(os.path.join(DST, "reftests"))
    for pkg in ("models", "configs", "data"):
        shutil.copytree(os.path.join(SRC, pkg), os.path.join(DST, pkg),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for t in ("test_kerple.py", "test_circulant_string.py", "test_performer.py"):
        shutil.copy2(os.path.join(SRC, t), os.path.join(DST, "reftests", t))
    manifest = {}
    for base, _, files in os.walk(DST):
        for f in sorted(files):
            p = os.path.join(base, f)
            manifest[os.path.relpath(p, DST)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    print(f"installed {len(manifest)} reference files into {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
