"""Builds the TEST-ONLY host simulation of the device templates (see hostsim.cpp)."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "paladin_b200", "csrc")


def build(tsan=False):
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhostsim_tsan.so" if tsan else "libhostsim.so")
    srcs = [os.path.join(HERE, "hostsim.cpp")]
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh"))
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps):
        return so
    flags = ["-O1", "-g", "-fsanitize=thread"] if tsan else ["-O2", "-g"]
    subprocess.check_call(["g++", "-std=c++17"] + flags + ["-shared", "-fPIC", "-o", so] + srcs + ["-lpthread"])
    return so


if __name__ == "__main__":
    print(build())
    print(build(tsan=True))
