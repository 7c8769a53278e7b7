"""TEST INFRASTRUCTURE ONLY -- runs one op of egotap_b200.capi.CudaBackend on CPU-resident test arguments: every
distinct CPU storage among the arguments is mirrored once on the GPU (so aliasing views stay aliased), the op runs
through the C ABI on the device, and the storages are copied back.  Lets the op-level tests written against the
CPU emulation run unchanged against the real kernels on the B200."""
import torch


class GpuOpAdapter:
    name = "cuda"

    def __init__(self, backend=None, to_device=None, synchronize=None):
        """defaults: the CUDA backend on the current device.  (tests/test_train_kernels.py also runs the adapter itself on
        the CPU -- emulation backend, `to_device` = clone -- so its mirroring / aliasing / copy-back logic is checked
        before it is trusted on the GPU box)"""
        if backend is None:
            from egotap_b200 import capi
            backend = capi.CudaBackend()
        self.be = backend
        self._to_device = to_device or (lambda t: t.cuda())
        self._sync = synchronize or torch.cuda.synchronize

    def __getattr__(self, op):
        fn = getattr(self.be, op)

        def call(*args):
            mirrors = {}

            def dev(t):
                st = t.untyped_storage()
                key = st.data_ptr()
                if key not in mirrors:
                    host = torch.empty(0, dtype=torch.uint8).set_(st)
                    mirrors[key] = (host, self._to_device(host))
                d8 = mirrors[key][1]
                typed = d8.view(t.dtype)
                return torch.as_strided(typed, t.size(), t.stride(), t.storage_offset())

            def conv(a):
                if isinstance(a, torch.Tensor):
                    return dev(a)
                if isinstance(a, (list, tuple)) and a and all(isinstance(x, torch.Tensor) for x in a):
                    return [dev(x) for x in a]
                return a
            out = fn(*[conv(a) for a in args])
            self._sync()
            for host, d8 in mirrors.values():
                host.copy_(d8.cpu())
            return out
        return call
