"""One launch of the deep-stack kernel (profiling target): python tools/deep_once.py [n_frames] [rows]"""
import sys, torch
sys.path.insert(0, ".")
import astroburst_b200 as ab
from astroburst_b200 import synth
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0"); ctx = ab.default_context(0)
fr = synth.stack_frames_torch(n_frames, rows, 4096, dev, seed=5)
cube = torch.stack(fr); del fr
out = torch.empty((rows, 4096), dtype=torch.float32, device=dev)
rej = torch.zeros(1, dtype=torch.int64, device=dev)
cfg = ab.StackConfig(3.0, 3.0, 5, False)
ptrs = [cube[f].data_ptr() for f in range(n_frames)]
for _ in range(2):
    ab.stack_images_device(ptrs, rows * 4096, cfg, out.data_ptr(), rej.data_ptr(),
                           torch.cuda.current_stream().cuda_stream, ctx)
torch.cuda.synchronize()
print("rejected", int(rej))
