#!/usr/bin/env python
"""Generates tests/golden/reference_py_vectors.json by IMPORTING the reference's own Python restatements of the
wave model (they run only in this container, /root/reference is absent on the GPU box):
  /root/reference/scripts/reflections/fresnel.py        fresnel_reflect_dir, fresnel_refract_dir (:25-57),
                                                        incident_angle / transmission_angle (:9-13)
  /root/reference/scripts/maxwell_boltzmann.py          maxwell_boltzmann_pdf, _a_from_mode (:6-10)
matplotlib is not installed here and is only used by the scripts' __main__ blocks, so it is stubbed.
The rs/rp/Reff formulas of fresnel.py live inside its `render()` closure (:125-148); they are evaluated here by
the same expressions with the script's eps replaced by the C++ value 1e-4 (radar_algorithms.h:110) — flagged
"restated" in the JSON, unlike the "imported" direction vectors.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference/scripts"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_py_vectors.json")


def _stub_matplotlib():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.widgets", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        m = types.ModuleType(name)
        m.__dict__.setdefault("__path__", [])
        sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib.widgets"].Slider = object
    sys.modules["matplotlib.widgets"].Button = object
    sys.modules["mpl_toolkits"].mplot3d = sys.modules["mpl_toolkits.mplot3d"]


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    _stub_matplotlib()
    fr = _load(os.path.join(REF, "reflections", "fresnel.py"), "ref_fresnel")
    mb = _load(os.path.join(REF, "maxwell_boltzmann.py"), "ref_mb")
    rng = np.random.default_rng(20240310)
    cases = []
    normal = np.array([0.0, 1.0])
    for k in range(64):
        ang = float(rng.uniform(0.0, 89.0)) * np.pi / 180.0
        n1 = float(rng.choice([1.0, 0.3, 0.03, 1.3, 2.5]))
        n2 = float(rng.choice([1.0, 0.3, 0.03, 1.3, 2.5, 10.0]))
        ray = np.array([np.sin(ang), -np.cos(ang)])
        refl, _ = fr.fresnel_reflect_dir(normal, ray, n1, n2)
        refr, _ = fr.fresnel_refract_dir(normal, ray, n1, n2)
        ia = float(fr.incident_angle(normal, ray))
        with np.errstate(all="ignore"):
            ta = float(fr.transmission_angle(normal, refr))
        eps = 0.0001
        if ia + ta < eps:
            rs = (n1 - n2) / (n1 + n2)
            rp = rs
        elif ia + ta > np.pi - eps:
            rs = rp = 1.0
        else:
            rs = -np.sin(ia - ta) / np.sin(ia + ta)
            rp = np.tan(ia - ta) / np.tan(ia + ta)
        reff = 0.5 * (rs * rs + rp * rp)
        cases.append({"angle": ang, "n1": n1, "n2": n2, "ray": ray.tolist(), "reflect_dir": np.asarray(refl).tolist(),
                      "refract_dir": np.asarray(refr).tolist(), "incident_angle": ia, "transmission_angle": ta,
                      "Reff_restated": float(reff)})
    mbv = []
    for mode in (1.0, 5.0, 17.0, 20.0):
        a = float(mb.maxwell_boltzmann_a_from_mode(mode))
        for x in (0.0, 1.0, 3.0, 10.0, 17.0, 33.0, 49.0):
            mbv.append({"mode": mode, "x": x, "pdf": float(mb.maxwell_boltzmann_pdf(x, a))})
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    json.dump({"source": "imported from /root/reference/scripts/{reflections/fresnel.py,maxwell_boltzmann.py}",
               "fresnel": cases, "maxwell_boltzmann": mbv}, open(OUT, "w"), indent=1)
    print("wrote", OUT, len(cases), len(mbv))


if __name__ == "__main__":
    main()
