"""TEST INFRASTRUCTURE ONLY -- builds tests/cuda_emu/_build/libtrain_emu.so: the training kernels' own source files
compiled by g++ against the CUDA-on-CPU shim (cuda_emu.h), and a backend object that drives it."""
import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "egotap_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libtrain_emu.so")
SOURCES = [os.path.join(HERE, "cuda_emu.cpp"), os.path.join(CSRC, "train_ops.cu"), os.path.join(CSRC, "train_model.cu"),
           os.path.join(CSRC, "gt_heatmap.cu")]
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")


def build(force=False):
    deps = SOURCES + [os.path.join(HERE, "cuda_emu.h"), os.path.join(CSRC, "host_util.cuh"), os.path.join(CSRC, "numeric.cuh"),
                      os.path.join(ROOT, "include", "egotap_b200.h")]
    if not force and os.path.isfile(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O2", "-fPIC", "-shared", "-DEB_HOST_EMU", "-I", HERE, "-I", CUDA_INC, "-I", CSRC,
           "-o", OUT] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building the kernel emulation library")
    return OUT


def make_backend():
    """A backend for training.TrainEngine whose TRAINING ops execute the real kernel source on the CPU emulation, and
    whose round-1 ops (tcgen05 GEMM, fused attention, LayerNorm, ingest, head, packing kernels: GPU-verified, not
    emulatable) are served by the op oracle."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import op_oracle
    from egotap_b200 import capi

    lib = C.CDLL(build())
    lib.egotap_b200_last_error.restype = C.c_char_p
    lib.egotap_b200_launch_count.restype = C.c_longlong
    for name, args in capi._TRAIN_ARGTYPES(C.c_void_p, C.c_longlong, C.c_int, C.c_float).items():
        getattr(lib, name).argtypes = args

    def check(rc, what):
        if rc != 0:
            raise RuntimeError("emulated %s failed (%d): %s" % (what, rc, lib.egotap_b200_last_error().decode()))

    orc = op_oracle.OracleBackend()

    class EmuBackend(capi.CudaBackend):
        name = "emu"

        def __init__(self):                      # no CUDA device, no product library
            self.L = _Checked(lib, check)
            self.device = torch.device("cpu")
            self._preset_of_J = {15: 0, 17: 1}

        @staticmethod
        def _st():
            return None

        def empty(self, shape, dtype=None):
            return orc.empty(shape, dtype or torch.float32)

        # round-1 ops: oracle
        gemm = staticmethod(orc.gemm)
        split2d = staticmethod(orc.split2d)
        ingest = staticmethod(orc.ingest)
        fill_dummy = staticmethod(orc.fill_dummy)
        pos_permute = staticmethod(orc.pos_permute)
        layernorm = staticmethod(orc.layernorm)
        attention = staticmethod(orc.attention)
        pu_bridge_gate = staticmethod(orc.pu_bridge_gate)
        head = staticmethod(orc.head)
        add3 = staticmethod(orc.add3)

    return EmuBackend(), orc


class _Checked:
    """wraps the emulation library so that the product's check(rc, what) helper (which reads the PRODUCT library's error
    string) is bypassed: a non-zero return raises here with the emulation library's own message"""

    def __init__(self, lib, check):
        self._lib, self._check = lib, check

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if name in ("egotap_b200_launch_count", "egotap_b200_last_error"):
            return fn

        def call(*a):
            self._check(fn(*a), name)
            return 0
        return call


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
