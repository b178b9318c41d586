"""Generates tests/golden/feed.npz from the reference's own HSIdataset4PT (run in the build container only:
needs /root/reference; `timm`, which Model_Pretraining.py imports but the feed does not use, is stubbed).

    python oracle/make_golden_feed.py
"""
import os, random, sys, types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.dont_write_bytecode = True


def import_reference_pretraining():
    if "timm" not in sys.modules:   # absent in this image; only CosineLRScheduler is imported from it
        timm = types.ModuleType("timm"); sched = types.ModuleType("timm.scheduler")
        sched.CosineLRScheduler = object
        timm.scheduler = sched
        sys.modules["timm"] = timm; sys.modules["timm.scheduler"] = sched
    # the driver does `from Models import HSIMAE`: give it THIS repository's drop-in module (and keep the reference's
    # own Models.py out of sys.modules, where it would shadow ours for everything imported later)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import Models  # noqa: F401
    sys.path.insert(0, "/root/reference")
    try:
        import Model_Pretraining as MP
    finally:
        sys.path.remove("/root/reference")
    return MP


def synthetic(seed=0):
    rng = np.random.default_rng(seed)
    scenes = [rng.standard_normal((13, 17, 32)).astype(np.float32), (5.0 * rng.standard_normal((9, 30, 32)) + 2.0).astype(np.float32)]
    rows = []
    for num, s in enumerate(scenes):
        mx, mn = (1, 0) if num == 0 else (int(s.max()), int(s.min()))   # norm=False / norm=True (int16 truncation, Preprocessing.py:114)
        for h in range(0, s.shape[0] - 8, 2):
            for w in range(0, s.shape[1] - 8, 3):
                rows.append((0, h, w, num, mx, mn))
    return scenes, np.array(rows, dtype=np.int16)


def main():
    MP = import_reference_pretraining()
    scenes, cut = synthetic()
    out = {"cut_info": cut, "scene0": scenes[0], "scene1": scenes[1]}
    # per-sample outputs with the reference's own flip draws
    ds = MP.HSIdataset4PT([scenes, cut], train=True)
    random.seed(123)
    out["train_items"] = np.stack([ds[i].numpy() for i in range(len(ds))])
    ds_eval = MP.HSIdataset4PT([scenes, cut], train=False)
    out["eval_items"] = np.stack([ds_eval[i].numpy() for i in range(len(ds_eval))])
    # one seeded epoch through the reference's DataLoader (order + flips + collation)
    from torch.utils.data import DataLoader
    torch.manual_seed(7); random.seed(7)
    out["epoch_batches"] = np.concatenate([b.numpy() for b in DataLoader(ds, batch_size=5, shuffle=True, num_workers=0, pin_memory=False)])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "feed.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
