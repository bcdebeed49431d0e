"""Launcher with the command line of the reference's build/simulation_launcher.py:

    python simulation_launcher.py nodiff      # path A: imhd-cuda_nodiff, 49 values of input.inp
    python simulation_launcher.py <other>     # path B: imhd-cuda, 37 values of input_diffusion.inp

As in the reference (build/simulation_launcher.py:8-47) every data file in the output directory except README.md is
deleted first, and the VALUES of the key=value input file are passed positionally in file order (keys are ignored).
The reference reads one input.inp for both drivers although main.cu's 37-argument order differs from it
(src/on-device/README.md:7-8 "out of phase"); here the diffusion driver gets its own file in main.cu's order.
Optional: --input FILE, --data-dir DIR (overrides path_to_data), IMHD_OUTPUT_EVERY=n in the environment.
"""
import argparse
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(os.path.dirname(HERE), "bin")


def read_values(path):
    with open(path) as f:
        return [line.split("=", 1)[1].strip() for line in f if "=" in line]


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("mode")
    ap.add_argument("--input")
    ap.add_argument("--data-dir")
    a = ap.parse_args(argv)
    nodiff = a.mode == "nodiff"
    driver = os.path.join(BIN, "imhd-cuda_nodiff" if nodiff else "imhd-cuda")
    values = read_values(a.input or os.path.join(HERE, "input.inp" if nodiff else "input_diffusion.inp"))
    ipath = 14 if nodiff else 13  # 0-based position of path_to_data
    if a.data_dir:
        values[ipath] = os.path.join(a.data_dir, "")
    data_root = values[ipath]
    os.makedirs(data_root, exist_ok=True)
    for name in os.listdir(data_root):
        full = os.path.join(data_root, name)
        if os.path.isfile(full) and name != "README.md":
            os.remove(full)
    return subprocess.run([driver] + values).returncode


if __name__ == "__main__":
    sys.exit(main())
