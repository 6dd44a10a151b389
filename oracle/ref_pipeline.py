"""The UNMODIFIED reference functions run as the CPU arm of bench.py (test infrastructure).

``install(reference_root)`` — called by ``__graft_entry__.build()`` in the build container, where
``/root/reference`` is mounted — places the reference's own ``semantic-slam`` Python tree under
``baseline/_ref/`` (git-ignored, so it never enters the repository history; not gpurun-ignored, so
it travels to the GPU box, which has no ``/root/reference``).  ``load()`` imports it from there with
stub ``timm`` / ``matplotlib`` modules (absent from the image; the ViT and the plotting code are never
called).  ``run_sequence`` then times exactly the reference's per-frame and per-pair calls
(SURVEY.md §8(d) "pipeline P"): ``KeypointSelector.select_keypoints`` -> ``DinoBackbone.pixel_to_patch``
-> ``DinoBackbone.extract_at_keypoints`` -> ``DescriptorRefiner.forward`` per frame and
``MatchVisualizer.find_matches`` per consecutive pair, fanned out over a process pool (the reference
itself is single-process Python; the pool lets the CPU arm use every host core).
"""

import os
import shutil
import sys
import time
import types
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
_W = {}


def install(reference_root="/root/reference"):
    """Copy <reference>/semantic-slam (Python files only) to baseline/_ref/semantic-slam."""
    src = os.path.join(reference_root, "semantic-slam")
    if not os.path.isdir(src):
        return False
    dst = os.path.join(REF_DIR, "semantic-slam")
    if os.path.isdir(dst):
        shutil.rmtree(dst)

    def ignore(d, names):
        return [n for n in names if not (os.path.isdir(os.path.join(d, n)) or n.endswith(".py"))
                or n in ("__pycache__", "data", "checkpoints", "wandb")]
    shutil.copytree(src, dst, ignore=ignore)
    return True


def available():
    return os.path.isfile(os.path.join(REF_DIR, "semantic-slam", "models", "keypoint_selector.py"))


def load():
    """Import the reference modules from baseline/_ref (never from /root/reference at run time)."""
    if "ref" in _W:
        return _W["ref"]
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    sys.dont_write_bytecode = True
    for name in ("timm", "matplotlib", "matplotlib.pyplot", "matplotlib.patches"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib.patches"].Circle = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    pkg = os.path.join(REF_DIR, "semantic-slam")
    # the drop-in package uses the same top-level module names ("models", "visualize_matches"):
    # the reference must win in this process
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "visualize_matches"]:
        del sys.modules[k]
    # (the reference's `models` has no __init__.py: as a namespace package it would lose against the
    # drop-in's regular package wherever that sits on sys.path, so that entry is hidden meanwhile)
    saved = list(sys.path)
    sys.path[:] = [pkg] + [p for p in saved if not os.path.isfile(os.path.join(p or ".", "models", "__init__.py"))]
    try:
        from models.keypoint_selector import KeypointSelector
        from models.dino_backbone import DinoBackbone
        from models.descriptor_refiner import DescriptorRefiner
        from visualize_matches import MatchVisualizer
    finally:
        sys.path[:] = saved
    _W["ref"] = dict(KeypointSelector=KeypointSelector, DinoBackbone=DinoBackbone,
                     DescriptorRefiner=DescriptorRefiner, MatchVisualizer=MatchVisualizer)
    return _W["ref"]


def _init_worker(state_dict, dims):
    import torch
    torch.set_num_threads(1)
    ref = load()
    C, Hd, D, layers = dims
    m = ref["DescriptorRefiner"](C, Hd, D, layers).eval()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in state_dict.items()})
    _W["refiner"] = m
    _W["selector"] = ref["KeypointSelector"](8, 8)          # select_keypoints uses no weights
    _W["ns"] = types.SimpleNamespace(patch_size=16)


def _extract_job(args):
    import torch
    sal, feat, K = args
    ref = load()
    with torch.no_grad():
        kp, sc = _W["selector"].select_keypoints(torch.from_numpy(sal)[None, :, :, None], num_keypoints=K)
        pc = ref["DinoBackbone"].pixel_to_patch(_W["ns"], kp)
        f = ref["DinoBackbone"].extract_at_keypoints(None, torch.from_numpy(feat)[None], pc)
        d = _W["refiner"](f)
    return kp[0].numpy(), sc[0].numpy(), d[0].numpy()


def _match_job(args):
    from .pipeline import pair_record
    d1, d2, kp1, kp2 = args
    m = load()["MatchVisualizer"].find_matches(None, d1, d2, 0.8)
    return pair_record([(i, j) for i, j, _ in m], kp1, kp2)


def run_sequence(sal, feat, state_dict, dims, K, workers):
    """sal (T,H,W) fp32, feat (T,h,w,C) fp32 NumPy -> ([pipeline.pair_record per consecutive pair], seconds)."""
    T = sal.shape[0]
    with ProcessPoolExecutor(workers, initializer=_init_worker, initargs=(state_dict, dims)) as pool:
        e = np.eye(4, 8, dtype=np.float32)
        list(pool.map(_match_job, [(e, e, e[:, :2], e[:, :2])] * workers))
        t0 = time.perf_counter()                      # pool start-up / imports are not the algorithm
        ex = list(pool.map(_extract_job, [(sal[t], feat[t], K) for t in range(T)]))
        res = list(pool.map(_match_job, [(ex[t][2], ex[t + 1][2], ex[t][0], ex[t + 1][0]) for t in range(T - 1)]))
        sec = time.perf_counter() - t0
    return res, sec
