"""Host-buffer front end: frames in pinned host memory -> poses in pinned host memory.

This is the call a user with host-side data makes (the reference's loop does the same thing
implicitly: ``set_input`` copies each batch to the GPU, reference
model/egotap_autoencoder_model.py:155-174, and metrics read the pose back).  Uploads run on a
copy stream, double-buffered against the compute stream, so the 1.47 MB/frame H2D transfer of
batch i+1 overlaps the kernels of batch i."""
import torch


class HostPipeline:
    def __init__(self, net, batch, depth=2):
        self.net = net
        self.device = next(net.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("egotap_b200 has no CPU path: move the module to a CUDA device first")
        self.depth = depth
        shape = (batch, net.channels_heatmap, net.W, net.H)
        self.dev_in = [torch.empty(shape, device=self.device) for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.free = [torch.cuda.Event() for _ in range(depth)]
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.h2d_bytes_per_step = self.dev_in[0].numel() * 4
        self.d2h_bytes_per_step = batch * net.num_joints * 3 * 4

    def run(self, host_batches, host_out):
        """host_batches: iterable of pinned fp32 CPU tensors (B, 6J, 64, 64); host_out: list of pinned
        (B, num_joints, 3) tensors receiving the poses.  Returns after everything has landed."""
        compute = torch.cuda.current_stream(self.device)
        for ev in self.free:
            ev.record(compute)
        for i, hb in enumerate(host_batches):
            slot = i % self.depth
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self.free[slot])
                self.dev_in[slot].copy_(hb, non_blocking=True)
                self.ready[slot].record(self.copy_stream)
            compute.wait_event(self.ready[slot])
            pose = self.net.predict_pose(self.dev_in[slot])
            self.free[slot].record(compute)
            host_out[i].copy_(pose, non_blocking=True)
        compute.synchronize()
        return host_out
