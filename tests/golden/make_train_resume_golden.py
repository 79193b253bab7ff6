"""Generates tests/golden/train_resume_ref.json by executing the reference's OWN statements for
  * resolving `training_config.resume_from_checkpoint` (train_denoiser.py:348-374: "latest", a path whose basename is
    looked up under `output_dir`, nothing to resume from) and
  * pruning old checkpoints before a save (`checkpoints_total_limit`, train_denoiser.py:1195-1225)
on directory trees made here.  tests/test_host_cpu.py requires the same answers of train_denoiser.resolve_resume_checkpoint /
prune_checkpoints in this repo.  Run here (needs /root/reference):  python tests/golden/make_train_resume_golden.py"""
import ast
import json
import os
import shutil
import sys
import tempfile
import types
from pathlib import Path

sys.path.insert(0, str(Path(__file__).parent))
from make_train_pack_golden import statements  # noqa: E402

REF = Path("/root/reference/train_denoiser.py")

# (directories present under output_dir, resume_from_checkpoint)
RESUME_CASES = [
    (["checkpoint-500", "checkpoint-1500", "checkpoint-1000", "logs"], "latest"),
    (["checkpoint-500", "checkpoint-1500"], "checkpoint-500"),
    (["checkpoint-500"], "/somewhere/else/checkpoint-500"),
    (["logs"], "latest"),
    ([], None),
    ([], ""),
]
# (directories present, checkpoints_total_limit)
PRUNE_CASES = [
    (["checkpoint-100", "checkpoint-300", "checkpoint-200", "logs"], 3),
    (["checkpoint-100", "checkpoint-300", "checkpoint-200"], 1),
    (["checkpoint-100", "checkpoint-200"], 3),
    (["checkpoint-100", "checkpoint-200"], None),
    (["checkpoint-1000", "checkpoint-200", "checkpoint-30"], 2),
]


def tree_of(names):
    d = tempfile.mkdtemp()
    for n in names:
        os.makedirs(os.path.join(d, n))
    return d


def main():
    tree = ast.parse(REF.read_text())
    resume = compile(ast.Module(body=statements(tree, 348, 374), type_ignores=[]), str(REF), "exec")
    prune = compile(ast.Module(body=statements(tree, 1195, 1225), type_ignores=[]), str(REF), "exec")
    out = dict(resume=[], prune=[])
    for names, want in RESUME_CASES:
        d = tree_of(names)
        said = []
        args = types.SimpleNamespace(training_config=types.SimpleNamespace(resume_from_checkpoint=want, output_dir=d))
        ns = dict(os=os, args=args, accelerator=types.SimpleNamespace(print=lambda *a: said.append(" ".join(map(str, a)))))
        exec(resume, ns)
        path = ns["resume_checkpoint_path"]
        out["resume"].append(dict(dirs=names, resume_from_checkpoint=want,
                                  chosen=None if path is None else os.path.relpath(path, d),
                                  initial_global_step=ns["initial_global_step"],
                                  said=[s.replace(d, "<out>") for s in said]))
        shutil.rmtree(d)
    for names, limit in PRUNE_CASES:
        d = tree_of(names)
        said = []
        args = types.SimpleNamespace(training_config=types.SimpleNamespace(checkpoints_total_limit=limit, output_dir=d))
        ns = dict(os=os, shutil=shutil, args=args,
                  accelerator=types.SimpleNamespace(print=lambda *a: said.append(" ".join(map(str, a)))))
        exec(prune, ns)
        out["prune"].append(dict(dirs=names, limit=limit, left=sorted(os.listdir(d)), said=said))
        shutil.rmtree(d)
    Path(__file__).with_name("train_resume_ref.json").write_text(json.dumps(out, indent=1))
    for k, v in out.items():
        for c in v:
            print(k, c)


if __name__ == "__main__":
    main()
