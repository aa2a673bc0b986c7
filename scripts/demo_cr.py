#!/usr/bin/env python
"""Closed-loop demo over a directory of CommonRoad scenarios (reference: scripts/demo_cr.py).

    python scripts/demo_cr.py [--cfg_file cfgs/demo_config.yaml]
"""
import argparse
import os
import sys

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from fiss_plus_planner_b200.planners.benchmark.planning import planning  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="Demo")
    ap.add_argument("--cfg_file", type=str, default=os.path.join(os.getcwd(), "cfgs/demo_config.yaml"),
                    help="specify the config file for the demo")
    args = ap.parse_args()
    with open(args.cfg_file, "r") as f:
        cfg = yaml.safe_load(f)
    output_dir = os.path.join(os.getcwd(), cfg["OUTPUT_DIR"])
    input_dir = os.path.join(os.getcwd(), cfg["INPUT_DIR"])
    files = cfg["FILES"] if cfg["FILES"] else sorted(os.listdir(input_dir))
    for file in files:
        planning(cfg, output_dir, input_dir, file)
