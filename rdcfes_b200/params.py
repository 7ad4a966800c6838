"""Model parameters: the reference's `input.dat` keys -> the flat vectors of include/rdc.h.

Each table lists, in the order of the `es.parameters` reads inside the reference's assemble callback, the
GetPot key and the default that the model's `input()` gives it:
  ADPM    adpm.C:163-224 (defaults), adpm.C:364-414 (read order)
  PIHNA   pihna.C:192-233,            pihna.C:360-381
  RIPF    ripf.C:173-249,             ripf.C:379-408 and 699-702
  PROTEAS proteas.C:185-216,          proteas.C:378-410
  HCC     coupled_hcc.C:351-369,      coupled_hcc.C:452-461
Unknown keys are ignored exactly like GetPot does.
"""
from __future__ import annotations

import math
import re
from typing import Dict, List, Tuple

import numpy as np

ADPM, PIHNA, RIPF, PROTEAS, HCC = 0, 1, 2, 3, 4
MODEL_NAMES = {"adpm": ADPM, "pihna": PIHNA, "ripf": RIPF, "proteas": PROTEAS, "hcc": HCC}
NVARS = {ADPM: 3, PIHNA: 5, RIPF: 3, PROTEAS: 5, HCC: 3}
VAR_NAMES = {
    ADPM: ("PrP", "A_b", "Tau"),                   # adpm.C:26-28
    PIHNA: ("n", "c", "h", "v", "a"),              # pihna.C:30-34
    RIPF: ("HU", "cc", "fb"),                      # ripf.C:24-26
    PROTEAS: ("hos", "tum", "nec", "vsc", "oed"),  # proteas.C:29-33
    HCC: ("l", "c", "n"),                          # coupled_hcc.C:33-35
}


def _pulse(key, cm=0.0):
    return [(key, cm), (key + "/pulse/0", -1.0e-20), (key + "/pulse/1", +1.0e+20)]


def _sigmoid(key):
    return [(key, 0.0), (key + "/sigmoid/0", +1.0e+20), (key + "/sigmoid/1", +1.1e+20)]


def _trapezoid(key):
    return [(key, 0.0), (key + "/trapezoid/0", -1.1e-20), (key + "/trapezoid/1", -1.0e-20),
            (key + "/trapezoid/2", +1.0e+20), (key + "/trapezoid/3", +1.1e+20)]


_DEG = "deg2rad"  # marker: value in input.dat is degrees (adpm.C:192,212)

TABLES: Dict[int, List[Tuple[str, float]]] = {
    ADPM: ([("decay/PrP/time_exponent", 0.0)] + _pulse("decay/PrP")
           + _pulse("diffuse/A_b") + _pulse("taxis_1/A_b") + _pulse("taxis_2/A_b") + _sigmoid("produce/A_b")
           + _trapezoid("transform/A_b") + _pulse("decay/A_b")
           + _pulse("diffuse/Tau") + _pulse("taxis_1/Tau") + _pulse("taxis_2/Tau") + _sigmoid("produce/Tau")
           + _trapezoid("transform/Tau") + _pulse("decay/Tau")
           + [("taxis/A_b/angle", 89.9), ("taxis/Tau/angle", 89.9)]),
    PIHNA: [("cells_min_capacity", 0.0), ("cells_max_capacity", 1.0), ("cytokines_max_capacity", 1.0),
            ("cells_max_capacity/exponent", 1.0),
            ("necrosis/c", 0.0), ("necrosis/h", 0.0), ("necrosis/v", 0.0),
            ("diffuse/c", 0.0), ("taxis/c", 0.0), ("diffuse/h", 0.0), ("taxis/h", 0.0),
            ("produce/c", 0.0), ("switch/c/to/h", 0.0), ("switch/h/to/c", 0.0), ("switch/h/to/n", 0.0),
            ("diffuse/v", 0.0), ("taxis/v", 0.0), ("produce/v", 0.0),
            ("secrete/a/from/c", 0.0), ("secrete/a/from/h", 0.0), ("uptake/a/from/v", 0.0), ("decay/a", 0.0)],
    RIPF: [("volume_fraction/stroma", 0.0), ("volume_fraction/parenchyma", 0.0), ("volume_fraction/exponent", 1.0),
           ("volume_fraction/min_vacant", 1.0e-12), ("volume_fraction/max_vacant", float("nan")),  # 1 - min_vacant
           ("HU/phi/cc/build", 0.0), ("HU/phi/cc/decay", 0.0), ("HU/phi/cc/rate", 0.0),
           ("HU/phi/fb/build", 0.0), ("HU/phi/fb/decay", 0.0), ("HU/phi/fb/rate", 0.0), ("HU/phi/tolerance", 0.0),
           ("cc/kappa", 0.0), ("cc/kappa/RT/c", 0.0), ("cc/delta", 0.0), ("cc/delta/RT/a", 0.0),
           ("cc/delta/RT/b", 0.0),
           ("fb/lambda", 0.0), ("fb/lambda/RT/r", 0.0), ("fb/lambda/HU/r", -1.0),
           ("fb/omicro", 0.0), ("fb/omicro/RT/r", 0.0), ("fb/omicro/fb/b", 0.0),
           ("fb/omega", 0.0), ("fb/diffusion", 0.0), ("fb/haptotaxis", 0.0), ("fb/radiotaxis", 0.0),
           ("HU/min", -1000.0), ("HU/max", +1000.0),
           ("RT_dose/broad/fractions", 1.0), ("RT_dose/focus/fractions", 1.0)],
    PROTEAS: [(k, 1.0) for k in (
        "cells/total_capacity", "radiotherapy/max_dosage",
        "host/proliferation", "host/vsc_threshold", "host/RT_death_rate", "host/RT_exp_a", "host/RT_exp_b",
        "host/necrosis_rate",
        "tumour/diffusion", "tumour/diffusion_host", "tumour/proliferation", "tumour/vsc_threshold",
        "tumour/RT_death_rate", "tumour/RT_exp_a", "tumour/RT_exp_b", "tumour/necrosis_rate",
        "necrosis/clearance", "necrosis/slope", "necrosis/vsc_threshold",
        "vascular/proliferation", "vascular/necrosis_rate",
        "oedema/diffusion", "oedema/proliferation", "oedema/vsc_threshold", "oedema/RT_coeff", "oedema/RT_exp",
        "oedema/reabsorption_rate")],
    HCC: [("cells/min_capacity", 0.0), ("cells/max_capacity", 1.0), ("cells/max_capacity/exponent", 1.0),
          ("produce/l", 0.0), ("diffuse/c", 0.0), ("mechano/c", 0.0), ("produce/c", 0.0),
          ("necrosis/l", 0.0), ("necrosis/c", 0.0), ("necrosis/pressure", 0.0)],
}

_ANGLE_KEYS = {"taxis/A_b/angle", "taxis/Tau/angle"}


def parse_getpot(text: str) -> Dict[str, str]:
    """Minimal GetPot reader: `key = value` lines, `#` comments, optional quotes."""
    out: Dict[str, str] = {}
    for line in text.splitlines():
        line = line.split("#", 1)[0].strip()
        m = re.match(r"^([^=\s]+)\s*=\s*(.*)$", line)
        if m:
            out[m.group(1)] = m.group(2).strip().strip("'\"").strip()
    return out


def flat_params(model: int, values: Dict[str, float] | None = None) -> np.ndarray:
    """Flat parameter vector for `model`; `values` overrides the reference defaults by GetPot key."""
    values = dict(values or {})
    out = np.zeros(len(TABLES[model]))
    for k, (key, default) in enumerate(TABLES[model]):
        v = float(values.get(key, default))
        if key in _ANGLE_KEYS:
            v = v * (math.pi / 180.0)  # utils.h:79 degrees_to_radians
        out[k] = v
    if model == RIPF:
        keys = [k for k, _ in TABLES[RIPF]]
        i_min, i_max = keys.index("volume_fraction/min_vacant"), keys.index("volume_fraction/max_vacant")
        if "volume_fraction/max_vacant" not in values:
            out[i_max] = 1.0 - out[i_min]  # ripf.C:181-182
    return out


def params_from_input(model: int, path: str) -> Tuple[np.ndarray, Dict[str, str]]:
    with open(path) as fh:
        kv = parse_getpot(fh.read())
    vals = {}
    for key, _ in TABLES[model]:
        if key in kv:
            vals[key] = float(kv[key])
    return flat_params(model, vals), kv
