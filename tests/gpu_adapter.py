"""TEST INFRASTRUCTURE ONLY -- runs one op of egotap_b200.capi.CudaBackend on CPU-resident test arguments: every
distinct CPU storage among the arguments is mirrored once on the GPU (so aliasing views stay aliased), the op runs
through the C ABI on the device, and the storages are copied back.  Lets the op-level tests written against the
CPU emulation run unchanged against the real kernels on the B200."""
import torch


class GpuOpAdapter:
    name = "cuda"

    def __init__(self):
        from egotap_b200 import capi
        self.be = capi.CudaBackend()

    def __getattr__(self, op):
        fn = getattr(self.be, op)

        def call(*args):
            mirrors = {}

            def dev(t):
                st = t.untyped_storage()
                key = st.data_ptr()
                if key not in mirrors:
                    host = torch.empty(0, dtype=torch.uint8).set_(st)
                    mirrors[key] = (host, host.cuda())
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
            torch.cuda.synchronize()
            for host, d8 in mirrors.values():
                host.copy_(d8.cpu())
            return out
        return call
