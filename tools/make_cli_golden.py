#!/usr/bin/env python3
"""Generate tests/golden/cli_<case>.mp3: the byte stream the UNMODIFIED reference CLI (oracle/_ref/encode) writes
for the PCM stored in tests/golden/<case>.npz.  Run in the build container only (needs /root/reference compiled
by `make -C oracle ref`).  tests/test_gpu_dropin.py compares the output of oracle/_ref/encode_gpu (the reference's
own main() linked against libmp3gpu.so) with these files byte for byte."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import cli_flags, write_wav  # noqa: E402


def main():
    gd = os.path.join(ROOT, "tests", "golden")
    enc = os.path.join(ROOT, "oracle", "_ref", "encode")
    for f in sorted(os.listdir(gd)):
        if not f.endswith(".npz"):
            continue
        g = np.load(os.path.join(gd, f))
        with tempfile.TemporaryDirectory() as tmp:
            wav, mp3 = os.path.join(tmp, "in.wav"), os.path.join(tmp, "out.mp3")
            write_wav(wav, g["pcm"], int(g["sfreq"]))
            subprocess.run([enc] + cli_flags(g["pcm"].shape[0], int(g["sfreq"]), int(g["bitrate"])) + [wav, mp3], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            data = open(mp3, "rb").read()
        out = os.path.join(gd, "cli_" + f[:-4] + ".mp3")
        open(out, "wb").write(data)
        print(out, len(data), "bytes")


if __name__ == "__main__":
    main()
