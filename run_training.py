"""Drop-in for the reference's run_training.py: optional TensorBoard side process, then train.py as a child.
(The reference also opens a browser and needs `netifaces`; neither is part of the hot path.)"""
import os
import shutil
import subprocess
import sys


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    tb = None
    if shutil.which("tensorboard") and os.environ.get("PBT_TENSORBOARD", "0") == "1":
        tb = subprocess.Popen(["tensorboard", "--logdir", os.path.join("outputs", "logs"), "--host", "127.0.0.1"])
    try:
        return subprocess.run([sys.executable, os.path.join(here, "train.py"), *sys.argv[1:]]).returncode
    finally:
        if tb is not None:
            tb.terminate()


if __name__ == "__main__":
    sys.exit(main())
