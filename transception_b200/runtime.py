"""CUDA-graph runner for the inference forward.

One eager forward of the default model enqueues several hundred small kernels; replaying them as one CUDA graph
removes the per-launch host cost (streams and graphs instead of a tracing compiler).  The runner owns static
input/output buffers; ``run_host`` is the end-to-end call with pinned host buffers on both sides.
"""
import torch


class GraphRunner:
    def __init__(self, model, batch, in_ch=1, size=224, device="cuda", warmup=3):
        self.model = model.eval()
        self.device = torch.device(device)
        self.x = torch.zeros((batch, in_ch, size, size), device=self.device, dtype=torch.float32)
        self.stream = torch.cuda.Stream(device=self.device)
        self.graph = torch.cuda.CUDAGraph()
        from . import ops
        with torch.no_grad():
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                for _ in range(warmup):
                    self.model(self.x)
            self.stream.synchronize()
            n0 = ops.launches()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.y = self.model(self.x)
            self.kernels_per_replay = ops.launches() - n0
        torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def replay(self):
        """Enqueue one forward on the current stream (input = self.x, output = self.y)."""
        self.graph.replay()

    def run_device(self, x):
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.y

    def run_host(self, x_host, y_host):
        """x_host / y_host: pinned host tensors. Copies in, replays, copies out; caller synchronises."""
        self.x.copy_(x_host, non_blocking=True)
        self.graph.replay()
        y_host.copy_(self.y, non_blocking=True)
        return y_host
