"""Dual-branch fine-tuning step (BASELINE.json configs[3]): DualViT-Large, Salinas-shaped batches (32 labelled + 71
unlabelled 9x9x32 tiles, 17 classes, mask 0.8, lambda 10, drop_path 0.2), loss = lamda*loss_rec + CE as in
Model_Finetuning.py:150-166.  Prints one JSON line (steps/s, labelled+unlabelled patches/s)."""
import json, os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import Models

torch.manual_seed(0); random.seed(0)
model = Models.DualViT(img_size=9, patch_size=3, in_chans=1, bands=32, b_patch_size=8, num_class=17, embed_dim=256, depth=12, num_heads=16,
                       s_depth=9, drop_path=0.2, decoder_embed_dim=64, decoder_depth=8, decoder_num_heads=8, norm_pix_loss=True,
                       trunc_init=True).cuda().train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=5e-2)
crit = torch.nn.CrossEntropyLoss(ignore_index=0)
for nl, nu in ((32, 71), (1024, 2272)):
    x, xu = torch.randn(nl, 1, 32, 9, 9, device="cuda"), torch.randn(nu, 1, 32, 9, 9, device="cuda")
    y = torch.randint(1, 17, (nl,), device="cuda")
    def step():
        loss_rec, _, _, logits = model(x, xu, mask_ratio=0.8)
        loss = 10 * loss_rec + crit(logits, y)
        opt.zero_grad(); loss.backward(); opt.step()
        return loss
    for _ in range(5): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): l = step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print(json.dumps({"metric": "dual-branch fine-tuning step (DualViT-Large, fwd+bwd+AdamW)", "labelled": nl, "unlabelled": nu,
                      "ms_per_step": dt * 1e3, "patches_per_s": (nl + nu) / dt, "loss": float(l)}))
