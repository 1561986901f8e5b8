"""Builds tests/golden/config_census.json: the `neural_net`, `render` and `training` blocks (plus dataset_type) of every
YAML under the reference's config_files/, so that the GPU box can check that each of the 47 experiment definitions
constructs and steps in this framework without the reference checkout.   python tests/golden/make_config_census.py"""
import glob
import json
import os

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = {}
for path in sorted(glob.glob("/root/reference/config_files/*.yaml")):
    try:
        cfg = yaml.safe_load(open(path))
    except yaml.YAMLError as e:          # two files are not valid YAML in the reference itself (mis-indented keys)
        out[os.path.basename(path)] = {"unparseable": str(e).splitlines()[0]}
        continue
    out[os.path.basename(path)] = {k: cfg[k] for k in ("dataset_type", "neural_net", "render", "training")}
    out[os.path.basename(path)].update({k: cfg.get(k) for k in ("tasks_to_perform", "video")})
json.dump(out, open(os.path.join(ROOT, "tests/golden/config_census.json"), "w"), indent=0, sort_keys=True)
print(len(out), "configs")
