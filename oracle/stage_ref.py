"""Stage the UNMODIFIED reference env package into `oracle/_ref/` (git-ignored, but shipped to the GPU
box with the repo snapshot) so that bench.py can time the reference's own NumPy path -- SURVEY.md
8(d) tiers T1 / T2 / T3 -- on the GPU box's host cores, in the same run as the CUDA path.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Nothing is edited: the two files are copied byte for byte from
the read-only checkout and their SHA-256 digests are written beside them.  The copy never enters the
git history (`oracle/_ref/` is in .gitignore); without the checkout (e.g. on the GPU box) this is a
no-op that keeps whatever was staged before.

    python -m oracle.stage_ref
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE_ROOT = os.environ.get("Q1_REFERENCE_ROOT", "/root/reference")
STAGED_ROOT = os.path.join(HERE, "_ref")
FILES = ("q1physrl_env/q1physrl_env/env.py", "q1physrl_env/q1physrl_env/phys.py", "LICENSE")


def stage() -> str:
    """Copy the reference env package into oracle/_ref/; returns the staged root ('' if there is
    neither a checkout nor an earlier staging)."""
    if not os.path.isfile(os.path.join(SOURCE_ROOT, FILES[0])):
        return STAGED_ROOT if os.path.isfile(os.path.join(STAGED_ROOT, FILES[0])) else ""
    digests = {}
    for rel in FILES:
        src, dst = os.path.join(SOURCE_ROOT, rel), os.path.join(STAGED_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            digests[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(STAGED_ROOT, "MANIFEST.json"), "w") as f:
        json.dump({"source": SOURCE_ROOT, "sha256": digests,
                   "note": "byte-for-byte copies of the reference's files, staged by oracle/stage_ref.py; "
                           "not part of this repository's history"}, f, indent=1)
    return STAGED_ROOT


if __name__ == "__main__":
    print(stage() or "no reference checkout and nothing staged")
