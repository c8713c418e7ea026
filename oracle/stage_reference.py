"""Stage the reference modules of the hot path into oracle/_ref/ (git-ignored, NOT gpurun-ignored).

TEST / BENCH INFRASTRUCTURE ONLY.  The reference is pure Python and cannot be pip-installed (no setup.py); /root/reference
does not exist on the GPU box.  `bench.py --impl reference` and the `cpu_baseline` leg must nevertheless time the
reference's OWN implementation on the box's host cores, so this recipe - run by `__graft_entry__.build()` in the
authoring container, where /root/reference is mounted - copies the handful of unmodified files the path imports into
`oracle/_ref/`, which travels to the box with the snapshot like a built .so.  Nothing is copied into git history
(`oracle/_ref/` is listed in .gitignore); `oracle/ref_shim.py` imports from it when /root/reference is absent.

    python oracle/stage_reference.py            # no-op when /root/reference is absent
"""
import os
import shutil

SRC = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
# the scoring path: backbone (models/timesformer.py + its two helpers), DINOHead (vision_transformer.py ->
# utils/utils.py:trunc_normal_), the scoring DINOLoss / export_loss (dino_similarity.py) and the loader's view builder
# (datasets_custom/dino_loss_loader.py -> transform.py, data_utils.py)
FILES = [
    "models/timesformer.py", "models/helpers.py", "models/vit_utils.py", "vision_transformer.py", "utils/utils.py",
    "dino_similarity.py", "datasets_custom/dino_loss_loader.py", "datasets_custom/transform.py",
    "datasets_custom/data_utils.py",
]


def stage():
    if not os.path.isfile(os.path.join(SRC, "models", "timesformer.py")):
        return False
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        src = os.path.join(SRC, rel)
        if not os.path.exists(dst) or os.path.getmtime(src) > os.path.getmtime(dst) or \
                os.path.getsize(src) != os.path.getsize(dst):
            shutil.copyfile(src, dst)
    return True


if __name__ == "__main__":
    print("staged" if stage() else "reference not mounted: nothing staged")
