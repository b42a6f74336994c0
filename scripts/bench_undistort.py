"""frame staging from a RAW (distorted) 752x480 camera image to the 640x480 pyramids,
(a) the reference's order through the ABI: undistort on the device, image back to the host, Frame construction (second H2D),
(b) fused: one H2D of the raw image, the remap feeds the pyramid kernel.  CUDA-event times per frame, median."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
from lsd_slam_b200 import abi

FOV = [0.535719308086809, 0.669566858850269, 0.493248545285398, 0.500408664348414, 0.897966326944875]
u = abi.UndistorterPTAM(FOV, (752, 480), "crop", (640, 480))
ctx = abi.Context(640, 480, u.getK(), max_frames=4)
u.install(ctx)
rng = np.random.default_rng(0)
raw = torch.from_numpy(rng.integers(0, 256, (480, 752)).astype(np.uint8)).pin_memory().numpy()
out = {}
for name in ("separate", "fused"):
    ts = []
    for i in range(60):
        ctx.timer_begin(0)
        if name == "separate":
            img = ctx.undistort(raw)
            ctx.upload(1, img)
        else:
            ctx.upload_distorted(1, raw)
        ctx.timer_end(0)
        ts.append(ctx.timer_ms(0) * 1e3)
        ctx.release(1)
    out[name + "_us_per_frame"] = float(np.median(ts[10:]))
print(json.dumps({"workload": "752x480 raw u8 -> undistort (crop) -> 640x480 pyramids + gradients + maxGradients", **out}))
