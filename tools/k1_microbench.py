"""Which part of K1 (feature fuse) costs what: full modality set vs the wide cosine modality alone vs the small
modalities alone (fp32-output entry point; relative numbers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tag_b200 as tb

DEV = "cuda:0"
vb = tb.make_videos(2500, 64, seed=3, device=DEV)
wv, ws, seg = tb.window_table([64] * 2500, 32, 8)
wv = torch.from_numpy(wv).to(DEV); ws = torch.from_numpy(ws).to(DEV)
full_r, full_d = tb.dims_maps(False)
cfgs = {"full": list(full_r), "vit only": ["vit"], "small only": ["global", "pose", "beta", "kp2d"], "kp2d only": ["kp2d"],
        "rot only": ["global", "pose"]}
for name, mods in cfgs.items():
    r = {m: full_r[m] for m in mods}; d = {m: full_d[m] for m in mods}
    fu = tb.FeatureFuser(r, d, DEV)
    dv = tb.DeviceVideos(vb, mods, DEV)
    mean = torch.zeros(fu.D, device=DEV); std = torch.ones(fu.D, device=DEV)
    for _ in range(2):
        fu.fuse(dv, wv, ws, 32, mean, std)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        feats, _ = fu.fuse(dv, wv, ws, 32, mean, std)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    byt = wv.numel() * 32 * (sum(r.values()) * 4 + fu.D * 4)
    print(f"{name:12s} D={fu.D:5d}: {ms:7.3f} ms  {byt / ms / 1e6:8.1f} GB/s algorithmic", flush=True)
    del feats
