"""Generates tests/golden/config_schema_ref.json by IMPORTING the reference's own univa/training/configuration_denoise.py
and univa/eval/configuration_eval.py (plain dataclasses, no third-party imports) and recording, per config class, every field's name, annotation and default;
plus, for each stage yaml the reference ships under scripts/denoiser/, its key set per section and whether every key is a
field of the schema (OmegaConf's structured merge rejects unknown keys, and so does this repo's loader).
Run here (needs /root/reference):  python tests/golden/make_config_golden.py"""
import dataclasses
import importlib.util
import json
from pathlib import Path

import yaml

REF = Path("/root/reference")


def main():
    spec = importlib.util.spec_from_file_location("ref_configuration_denoise", REF / "univa/training/configuration_denoise.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {"classes": {}, "yamls": {}}
    for cls in ("TrainingConfig", "DatasetConfig", "ModelConfig"):
        fields = {}
        for f in dataclasses.fields(getattr(mod, cls)):
            default = None if f.default is dataclasses.MISSING else f.default
            fields[f.name] = {"type": str(f.type).replace("typing.", ""), "default": default,
                              "has_default": f.default is not dataclasses.MISSING}
        out["classes"][cls] = fields
    spec2 = importlib.util.spec_from_file_location("ref_configuration_eval", REF / "univa/eval/configuration_eval.py")
    mod2 = importlib.util.module_from_spec(spec2)
    spec2.loader.exec_module(mod2)
    out["classes"]["EvalConfig"] = {
        f.name: {"type": str(f.type).replace("typing.", ""), "default": None if f.default is dataclasses.MISSING else f.default,
                 "has_default": f.default is not dataclasses.MISSING} for f in dataclasses.fields(mod2.EvalConfig)}
    sections = {"training_config": "TrainingConfig", "dataset_config": "DatasetConfig", "model_config": "ModelConfig"}
    for y in sorted((REF / "scripts/denoiser").glob("*.yaml")):
        raw = yaml.safe_load(y.read_text())
        rec = {}
        for sec, cls in sections.items():
            keys = sorted((raw.get(sec) or {}).keys())
            rec[sec] = {"keys": keys, "unknown": [k for k in keys if k not in out["classes"][cls]],
                        "values": {k: v for k, v in (raw.get(sec) or {}).items() if isinstance(v, (int, float, bool, str, type(None)))}}
        out["yamls"][y.name] = rec
    Path(__file__).with_name("config_schema_ref.json").write_text(json.dumps(out, indent=1, sort_keys=True))
    for n, r in out["yamls"].items():
        print(n, {s: v["unknown"] for s, v in r.items()})


if __name__ == "__main__":
    main()
