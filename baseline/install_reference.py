"""Install the UNMODIFIED reference (ai4in/YOLO-Contour-Regression, /root/reference/ultralytics-main) into the
git-ignored baseline/_ref so that it travels to the GPU box with the repo snapshot:

    python baseline/install_reference.py

runs  pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy>
on a copy under /tmp (the build writes into the source tree; /root/reference is read-only).  --no-deps: the
reference's requirements name matplotlib / seaborn / thop, which this image does not have; the run-time shims in
baseline/refload.py cover what the hot path touches.  No reference source enters the git history."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/ultralytics-main"
DST = os.path.join(HERE, "_ref")


def main():
    if not os.path.isdir(SRC):
        print(f"{SRC} not found: the reference can only be installed in the build container")
        return 1
    tmp = tempfile.mkdtemp(prefix="ycr_ref_")
    work = os.path.join(tmp, "src")
    shutil.copytree(SRC, work, ignore=shutil.ignore_patterns("docs", "docker", "examples", "*.png", "*~"))
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
           "/opt/wheelhouse", "--target", DST, work]
    r = subprocess.run(cmd, capture_output=True, text=True)
    print(r.stdout[-2000:], r.stderr[-2000:])
    shutil.rmtree(tmp, ignore_errors=True)
    if r.returncode != 0:
        return r.returncode
    # the model yamls the flow tests use live at the repository root of the reference, outside the package
    for name in ("yolov8-seg.yaml",):
        src = os.path.join(SRC, name)
        if os.path.exists(src):
            shutil.copy(src, os.path.join(DST, name))
    print("installed into", DST, ":", sorted(os.listdir(DST))[:8])
    return 0


if __name__ == "__main__":
    sys.exit(main())
