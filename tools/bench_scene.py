"""Dense per-pixel classification of a Salinas-sized synthetic scene (BASELINE.json configs[4]): 512x217 pixels,
32 GWPCA bands, HSIViT-Large, on-device sliding windows.  Prints one JSON line (pixels/s)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import Models
from hsimae_b200.scene import classify_scene

torch.manual_seed(0)
vit = Models.HSIViT(img_size=9, patch_size=3, in_chans=1, bands=32, b_patch_size=8, num_class=17, embed_dim=256, depth=12,
                    num_heads=16, s_depth=9, trunc_init=True).cuda().eval()
scene = torch.randn(512, 217, 32, device="cuda")
for batch in (4096, 16384):
    classify_scene(vit, scene, batch=batch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        out = classify_scene(vit, scene, batch=batch)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(json.dumps({"metric": "dense scene classification pixels/s (HSIViT-Large, 512x217x32, 9x9 windows)", "value": 512 * 217 / dt,
                      "unit": "pixels/s", "seconds_per_scene": dt, "batch": batch, "reference_cpu": "247 patches/s on 8 cores (BASELINE.md section 2)"}))
