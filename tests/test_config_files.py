"""Readers of the reference's configuration files (radarays_ros_b200/config.py): the rosparam material file format of
config/mulran_kaist02.yaml / config/oru4.yaml and the `dynparam dump` format of cfg/mulran_kaist_dyncfg.yaml. Written
samples in the same formats; where /root/reference is present (this container) its own files are read too and the dyn-cfg
dump must equal the MULRAN_DYNCFG table this repo's bench and tests use."""
import os

import pytest

from radarays_ros_b200 import MULRAN_DYNCFG, RadarModelConfig
from radarays_ros_b200.config import load_dyncfg_yaml, load_params_yaml

PARAMS = """# material file as loaded by <rosparam command="load" .../>
materials:
  # air
  - velocity: 0.3
    ambient: 1.0
    diffuse: 0.0
    specular: 1.0
  # wall stone
  - velocity: 0.0
    ambient: 1.0
    diffuse: 0.0
    specular: 3000.0

material_id_air: 0

object_materials: [
  1, # 0: HallwayGround-mesh
  1  # 1: Building-mesh
]
"""

DYNCFG = """!!python/object/new:dynamic_reconfigure.encoding.Config
dictitems:
  ambient_noise: 2
  beam_width: 10.0
  energy_max: 0.72
  include_motion: false
  n_cells: 3424
  n_samples: 50
  groups: !!python/object/new:dynamic_reconfigure.encoding.Config
    dictitems:
      ambient_noise: 2
      id: 0
      name: Default
      parameters: !!python/object/new:dynamic_reconfigure.encoding.Config
        state: []
    state: []
  signal_denoising_triangular_width: 35
state: []
"""


def test_params_yaml(tmp_path):
    p = tmp_path / "materials.yaml"
    p.write_text(PARAMS)
    mats, obj, air = load_params_yaml(p)
    assert mats == [(0.3, 1.0, 0.0, 1.0), (0.0, 1.0, 0.0, 3000.0)] and obj == [1, 1] and air == 0
    (tmp_path / "empty.yaml").write_text("object_materials: [0]\n")
    with pytest.raises(ValueError):
        load_params_yaml(tmp_path / "empty.yaml")


def test_dyncfg_dump(tmp_path):
    p = tmp_path / "dyncfg.yaml"
    p.write_text(DYNCFG)
    cfg = load_dyncfg_yaml(p)
    assert cfg.ambient_noise == 2 and cfg.beam_width == 10.0 and cfg.energy_max == 0.72 and cfg.n_samples == 50
    assert cfg.include_motion == 0 and cfg.signal_denoising_triangular_width == 35
    assert cfg.resolution == RadarModelConfig().resolution          # untouched fields keep the cfg defaults


@pytest.mark.skipif(not os.path.isdir("/root/reference/cfg"), reason="reference tree not present")
def test_reference_files():
    cfg = load_dyncfg_yaml("/root/reference/cfg/mulran_kaist_dyncfg.yaml")
    want = RadarModelConfig(**MULRAN_DYNCFG)
    for k, v in cfg.to_dict().items():
        if k == "include_motion":
            continue                                                  # the bench table switches it off (static poses)
        assert v == getattr(want, k), "cfg/mulran_kaist_dyncfg.yaml: %s = %r, MULRAN_DYNCFG has %r" % (k, v, getattr(want, k))
    mats, obj, air = load_params_yaml("/root/reference/config/mulran_kaist02.yaml")
    assert len(mats) == 2 and air == 0 and obj[0] == 1 and mats[1] == (0.0, 1.0, 0.0, 3000.0)
    mats4, obj4, air4 = load_params_yaml("/root/reference/config/oru4.yaml")
    assert len(obj4) == 18 and max(obj4) < len(mats4) and air4 == 0
