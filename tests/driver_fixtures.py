"""Shared helpers for the closed-loop driver tests: the committed CommonRoad scenario fixtures
(tests/golden/scenario_*.xml.gz, re-written from the reference's data/demo by make_golden_driver.py)."""
import glob
import gzip
import os

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENARIOS = sorted(os.path.basename(p)[len("scenario_"):-len(".xml.gz")]
                   for p in glob.glob(os.path.join(GOLDEN_DIR, "scenario_*.xml.gz")))
METHODS = ("FOP", "FOP+", "FISS", "FISS+")


def unpack_scenario(name: str, tmp_path) -> str:
    out = os.path.join(str(tmp_path), name + ".xml")
    with gzip.open(os.path.join(GOLDEN_DIR, f"scenario_{name}.xml.gz"), "rb") as g, open(out, "wb") as f:
        f.write(g.read())
    return out


def driver_golden(method: str, name: str) -> str:
    return os.path.join(GOLDEN_DIR, f"driver_{method.replace('+', 'plus')}_{name}.npz")
