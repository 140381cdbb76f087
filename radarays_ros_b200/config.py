"""The reference's configuration files, read without ROS.

* `load_params_yaml(path)`  — what `<rosparam command="load" file="config/mulran_kaist02.yaml"/>` puts on the parameter
  server and `Radar::loadParams` reads back (src/radarays_ros/Radar.cpp:220-226, ros_helper.cpp
  `loadRadarMaterialsFromParameterServer`): `materials` (list of velocity / ambient / diffuse / specular),
  `object_materials`, `material_id_air`. The older per-field layout of config/oru4.yaml (`velocities`, `ambient`,
  `diffuse`, `specular` arrays) is accepted as well.
* `load_dyncfg_yaml(path)`  — a `rosrun dynamic_reconfigure dynparam dump` file (cfg/mulran_kaist_dyncfg.yaml, loaded by
  launch/mulran_sim.launch:32-33 with `dynparam load`): the `RadarModelConfig` the dynamic-reconfigure server would
  deliver to `Radar::updateDynCfg`. The dump carries python object tags (`dynamic_reconfigure.encoding.Config`) and a
  `groups` subtree; only the top-level `dictitems` that name a field of cfg/RadarModel.cfg are used.
"""
import yaml

from .types import RadarModelConfig, _CFG_FIELDS


class _Loader(yaml.SafeLoader):
    pass


def _any_object(loader, suffix, node):
    # `!!python/object/new:dynamic_reconfigure.encoding.Config {dictitems: {...}, state: [...]}` -> its dictitems
    if isinstance(node, yaml.MappingNode):
        m = loader.construct_mapping(node, deep=True)
        return m.get("dictitems", m)
    if isinstance(node, yaml.SequenceNode):
        return loader.construct_sequence(node, deep=True)
    return loader.construct_scalar(node)


_Loader.add_multi_constructor("tag:yaml.org,2002:python/", _any_object)


def load_params_yaml(path):
    """-> (materials [(velocity, ambient, diffuse, specular)], object_materials [int], material_id_air int)."""
    with open(path) as f:
        doc = yaml.load(f, Loader=_Loader) or {}
    mats = []
    for m in doc.get("materials", []):
        mats.append((float(m["velocity"]), float(m["ambient"]), float(m["diffuse"]), float(m["specular"])))
    if not mats and "velocities" in doc:
        # older layout of config/oru3.yaml / oru4.yaml: one array per field (the current ros_helper.cpp no longer reads it)
        n = len(doc["velocities"])
        cols = [doc.get(k, [0.0] * n) for k in ("velocities", "ambient", "diffuse", "specular")]
        if any(len(c) != n for c in cols):
            raise ValueError("%s: velocities / ambient / diffuse / specular differ in length" % path)
        mats = [tuple(float(c[i]) for c in cols) for i in range(n)]
    if not mats:
        raise ValueError("%s: no `materials` list" % path)
    obj = [int(x) for x in doc.get("object_materials", [])]
    air = int(doc.get("material_id_air", 0))
    return mats, obj, air


def load_dyncfg_yaml(path, base=None):
    """-> RadarModelConfig: `base` (default: the cfg/RadarModel.cfg defaults) with every dumped field applied."""
    with open(path) as f:
        doc = yaml.load(f, Loader=_Loader) or {}
    if isinstance(doc, dict) and "dictitems" in doc:
        doc = doc["dictitems"]
    cfg = (base.copy() if base is not None else RadarModelConfig())
    names = {n for n, _, _ in _CFG_FIELDS}
    upd = {}
    for k, v in doc.items():
        if k in names and not isinstance(v, (dict, list)):
            upd[k] = (int(v) if isinstance(v, bool) else v)
    return cfg.update(**upd)
