"""Stage the UNMODIFIED reference modules of the hot path under the git-ignored `baseline/_ref/`.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/add_oracle.py's header).  The reference is a set of Python scripts
without setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference` has nothing to install;
its "install" is a file copy of the packages the path imports:

    modeling/   (ADD.py, operations.py, aspp_train.py, decoder.py, genotypes.py, sync_batchnorm/, siblings)
    utils/metrics.py
    searched_arch/autodeeplab/   (the genotype / network-path .npy files eval.py loads)

`baseline/_ref/` is listed in .gitignore (never committed — no reference source enters the history) but not in
.gpurunignore, so the copy travels to the GPU box with the snapshot, where `bench.py --impl reference` and the
`cpu_baseline` leg time the reference's own `ADD.dynamic_inference` + `Evaluator` on the host cores
(`cpu_baseline.kind = "reference"`).  When `/root/reference` is absent (the GPU box) this is a no-op and whatever was
staged earlier is used; when nothing was ever staged, bench.py falls back to the oracle port (`kind = "port"`).

Run by `__graft_entry__.build()`; also `python oracle/stage_reference.py`."""
from __future__ import annotations

import shutil
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
DST = ROOT / "baseline" / "_ref"
WHAT = ["modeling", "utils/metrics.py", "utils/__init__.py", "searched_arch/autodeeplab"]


def staged() -> bool:
    return (DST / "modeling" / "ADD.py").exists() and (DST / "utils" / "metrics.py").exists()


def stage(force: bool = False) -> bool:
    """Copy the reference's hot-path packages into baseline/_ref.  Returns True when a staged copy exists afterwards."""
    if not REF.exists():
        return staged()
    if staged() and not force:
        return True
    for rel in WHAT:
        src, dst = REF / rel, DST / rel
        if not src.exists():
            if rel.endswith("__init__.py"):          # utils/ is a namespace package in the reference
                continue
            raise FileNotFoundError(src)
        dst.parent.mkdir(parents=True, exist_ok=True)
        if src.is_dir():
            shutil.copytree(src, dst, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            shutil.copy2(src, dst)
    (DST / "STAGED_FROM").write_text(f"{REF} (unmodified copy; see oracle/stage_reference.py)\n")
    return staged()


if __name__ == "__main__":
    print("staged" if stage(force=True) else "reference not available", DST)
