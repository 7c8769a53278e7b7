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
SOURCES = [os.path.join(HERE, "cuda_emu.cpp"), os.path.join(HERE, "selftest.cu"), os.path.join(CSRC, "kernels.cu"), os.path.join(CSRC, "gemm_launch.cu"),
           os.path.join(CSRC, "attention.cu"), os.path.join(CSRC, "attention_bwd.cu"), os.path.join(CSRC, "pu_chain.cu"), os.path.join(CSRC, "pu_chain_bwd.cu"), os.path.join(CSRC, "metrics.cu"),
           os.path.join(CSRC, "plan.cu"),
           os.path.join(CSRC, "train_ops.cu"),
           os.path.join(CSRC, "train_model.cu"),
           os.path.join(CSRC, "gt_heatmap.cu")]
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
CUDA_LIB = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "lib64")


def build(force=False):
    deps = SOURCES + [os.path.join(HERE, "cuda_emu.h"), os.path.join(HERE, "ptx_emu.h"), os.path.join(CSRC, "host_util.cuh"),
                      os.path.join(CSRC, "numeric.cuh"), os.path.join(CSRC, "ptx.cuh"), os.path.join(CSRC, "gemm.cuh"),
                      os.path.join(ROOT, "include", "egotap_b200.h")]
    if not force and os.path.isfile(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    # -fsanitize=alignment: every pointer dereference is checked against its type's alignment (float4 / uint4: 16 bytes,
    # float2 / uint2: 8, packed bf16x2 stores: 4) and aborts on a violation -- the one class of kernel bug that plain x86
    # execution would forgive and the GPU would not
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O3", "-mavx2", "-fno-strict-aliasing", "-fPIC", "-shared", "-DEB_HOST_EMU", "-fsanitize=alignment",
           "-fno-sanitize-recover=alignment", "-I", HERE, "-I", CUDA_INC, "-I", CSRC, "-o", OUT] + SOURCES + \
          ["-L", CUDA_LIB, "-Wl,-rpath," + CUDA_LIB, "-lcudart"]     # error-string / event symbols only; no device is touched
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building the kernel emulation library")
    return OUT


def make_backend(real_tensor_core=False, all_oracle=False, dry=False):
    """A backend for training.TrainEngine whose bandwidth-bound ops -- the training kernels and the round-1 kernels of
    csrc/kernels.cu (ingest, LayerNorm, head, packing) -- execute the real kernel source on the CPU emulation; the
    tensor-core kernels (tcgen05 GEMM, fused attention: GPU-verified in round 1, not emulatable) are served by the op
    oracle."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import op_oracle
    from egotap_b200 import capi

    lib = C.CDLL(build())
    lib.egotap_b200_last_error.restype = C.c_char_p
    lib.egotap_b200_launch_count.restype = C.c_longlong
    for name, args in capi._TRAIN_ARGTYPES(C.c_void_p, C.c_longlong, C.c_int, C.c_float).items():
        getattr(lib, name).argtypes = args
    P, I = C.c_void_p, C.c_int
    lib.egotap_b200_gemm.argtypes = [C.POINTER(capi.Gemm), C.c_void_p]
    lib.egotap_b200_gemm_variant_name.restype = C.c_char_p
    lib.egotap_b200_attention.argtypes = [P] * 6 + [I, I, P]
    lib.egotap_b200_ingest.argtypes = [P, I, I] + [P] * 5
    lib.egotap_b200_layernorm.argtypes = [P] * 3 + [C.c_longlong, I, I, C.c_float] + [P] * 4
    lib.egotap_b200_head.argtypes = [P, I] + [P] * 5 + [C.c_longlong, I, P, P]

    def check(rc, what):
        if rc != 0:
            raise RuntimeError("emulated %s failed (%d): %s" % (what, rc, lib.egotap_b200_last_error().decode()))

    orc = op_oracle.OracleBackend()

    class EmuBackend(capi.CudaBackend):
        name = "emu"

        def __init__(self):                      # no CUDA device, no product library
            self.L = lib
            self.device = torch.device("cpu")
            self._preset_of_J = {15: 0, 17: 1}

        @staticmethod
        def _st():
            return None

        def _raise(self, name, rc):
            check(rc, name)

        def empty(self, shape, dtype=None):
            return orc.empty(shape, dtype or torch.float32)

        def _py(self, fn, *a, **k):              # an oracle-served op, recorded on the tape like a library call
            fn(*a, **k)
            if self._tape is not None:
                self._tape.append((lambda: fn(*a, **k), (), None))

        # the tensor-core kernels cannot be emulated: oracle.  Everything else (round-1 bandwidth kernels of
        # csrc/kernels.cu included) runs the real kernel source through the inherited CudaBackend methods.
        def gemm(self, *a, **k):
            if real_tensor_core:
                self.gemm_tc(*a, **k)
            else:
                self._py(orc.gemm, *a, **k)

        def attention(self, *a):
            if real_tensor_core:
                self.attention_tc(*a)
            else:
                self._py(orc.attention, *a)

        def attention_lse(self, *a):
            if real_tensor_core:
                capi.CudaBackend.attention_lse(self, *a)
            else:
                self._py(orc.attention_lse, *a)

        def attention_bwd(self, *a):
            if real_tensor_core:
                capi.CudaBackend.attention_bwd(self, *a)
            else:
                self._py(orc.attention_bwd, *a)

        # ... but their SOURCE does run on the functional tcgen05 / TMA / mbarrier model (ptx_emu.h), for op-level tests
        def gemm_tc(self, *a, **k):
            saved, capi.require_cuda = capi.require_cuda, (lambda *t: None)
            try:
                d = capi.gemm_desc(*a, **k)
            finally:
                capi.require_cuda = saved
            self._c("egotap_b200_gemm", C.byref(d), None)

        def attention_tc(self, qk_hi, qk_lo, vt_hi, vt_lo, ctx_hi, ctx_lo, frames, precision):
            p = lambda t: None if t is None else t.data_ptr()
            self._c("egotap_b200_attention", p(qk_hi), p(qk_lo), p(vt_hi), p(vt_lo), p(ctx_hi), p(ctx_lo), frames, precision, None)

        def gemm_variant_names(self):
            return [lib.egotap_b200_gemm_variant_name(i).decode() for i in range(lib.egotap_b200_gemm_num_variants())]

    if all_oracle:
        # every op served by the oracle but still recorded call by call: for tests of the engine's record / replay logic
        # that do not need the (slow) fiber execution of the kernels
        def served(name):
            return lambda self, *a, **k: self._py(getattr(orc, name), *a, **k)
        ops = [n for n in dir(orc) if not n.startswith("_") and callable(getattr(orc, n)) and n not in ("empty", "zero", "copy")]
        EmuBackend = type("OracleTapeBackend", (EmuBackend,), {n: served(n) for n in ops})
        EmuBackend.zero = lambda self, t: self._py(orc.zero, t)
        EmuBackend.copy = lambda self, d, s_: self._py(orc.copy, d, s_)
    if dry:
        # for emu_set_dry_run(1): every op goes to the library (host-side checks, tensor maps, launch geometry), nothing executes;
        # buffers are untouched virtual memory, so the engine can be driven at BASELINE.json's full batch sizes
        EmuBackend = type("DryBackend", (EmuBackend,), dict(
            empty=lambda self, shape, dtype=None: torch.empty(shape, dtype=dtype or torch.float32),
            zero=lambda self, t: None, copy=lambda self, d, s_: None,
            gemm=lambda self, *a, **k: self.gemm_tc(*a, **k), attention=lambda self, *a: self.attention_tc(*a),
            attention_lse=lambda self, *a: capi.CudaBackend.attention_lse(self, *a),
            attention_bwd=lambda self, *a: capi.CudaBackend.attention_bwd(self, *a)))
    return EmuBackend(), orc


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
