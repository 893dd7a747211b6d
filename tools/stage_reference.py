"""Stage the reference's hot-path sources for the GPU box (which has no /root/reference).

    python tools/stage_reference.py            # /root/reference -> baseline/_ref  (git-ignored, NOT gpurun-ignored)

Copies the directories the hot path lives in (``utils/``, ``model/``, ``gfnet_configs/``: 300 KB of Python, no weights)
unmodified, so that tests and ``bench.py --impl reference`` can import the reference's own functions on the GPU box
exactly as they do here from /root/reference (``oracle.reference.load_reference``).  Nothing under baseline/_ref is
tracked by git or imported by the product package.  ``__graft_entry__.build()`` runs this when /root/reference exists.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("GFNET_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
PARTS = ("utils", "model", "gfnet_configs")


def stage(src=SRC, dst=DST):
    if not os.path.isdir(src):
        return False
    for part in PARTS:
        s, d = os.path.join(src, part), os.path.join(dst, part)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.pth", "*.ckpt"))
    with open(os.path.join(dst, "STAGED_FROM"), "w") as f:
        f.write(src + "\n")
    return True


if __name__ == "__main__":
    ok = stage()
    print(("staged %s -> %s" % (SRC, DST)) if ok else ("%s not found: nothing staged" % SRC))
    sys.exit(0 if ok else 1)
