#!/usr/bin/env python
"""Install the UNMODIFIED reference package (``diffsynth``, from /root/reference/animation) into the git-ignored
``baseline/_ref/`` so that it travels to the GPU box with the gpurun snapshot (``.gitignore`` lists it,
``.gpurunignore`` does not).  No reference source enters the repository history.

    python baseline/install_ref.py          # idempotent; a no-op where /root/reference does not exist

This is the offline install the bench contract names:
``pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy of animation/>``
(from a copy under /tmp because setuptools writes build/ and *.egg-info into the source tree and /root/reference is
read-only; --no-deps because imageio / peft / accelerate / modelscope / ftfy are absent from the wheelhouse —
``baseline/ref_loader.py`` stubs exactly those imports, none of which is on the DiT hot path).
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/animation"
TARGET = os.path.join(HERE, "_ref")


def installed() -> bool:
    return os.path.isfile(os.path.join(TARGET, "diffsynth", "pipelines", "wan_video.py"))


def install(force: bool = False) -> bool:
    """True if baseline/_ref holds the reference afterwards."""
    if installed() and not force:
        return True
    if not os.path.isdir(REF_SRC):
        return False
    tmp = tempfile.mkdtemp(prefix="fairygen_ref_src_")
    try:
        src = os.path.join(tmp, "animation")
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns("outputs", "*.sh", "__pycache__"))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        subprocess.run(cmd, check=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return installed()


if __name__ == "__main__":
    ok = install(force="--force" in sys.argv)
    print(f"baseline/_ref: {'installed' if ok else 'UNAVAILABLE (no /root/reference here and nothing installed)'}")
    sys.exit(0 if ok else 1)
