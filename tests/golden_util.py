"""Load tests/golden/*.json.gz and regenerate their inputs."""
import gzip
import hashlib
import json
from pathlib import Path

import numpy as np

import checker as C
from dump1090_b200 import synth

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"
NAMES = sorted(p.name[: -len(".json.gz")] for p in GOLDEN_DIR.glob("*.json.gz"))


def load(name):
    with gzip.open(GOLDEN_DIR / f"{name}.json.gz", "rb") as f:
        return json.loads(f.read().decode())


def make_input(doc) -> np.ndarray:
    if doc["generator"] == "modes1":
        data = C.modes1()
    else:
        data = getattr(synth, doc["generator"])(**doc["args"])
    assert data.size == doc["nbytes"]
    assert hashlib.sha256(data.tobytes()).hexdigest() == doc["sha256"], "golden input drifted"
    return data


def cases():
    for name in NAMES:
        doc = load(name)
        for cid, case in doc["cases"].items():
            yield name, cid


CASES = list(cases())
