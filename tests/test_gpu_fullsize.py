"""BASELINE.json configs[1] at FULL size on the GPU (urban-5M, 400 x 3360, 3 passes, 256 samples):
one frame against the CPU oracle bit for bit, plus size-independent properties on a pose batch
(determinism, batch == single, azimuth shards == full frame, frame-id keyed noise)."""
import ctypes as C

import numpy as np
import pytest
import torch

from radarays_ros_b200 import MULRAN_DYNCFG, RadarModelConfig, scenes
from radarays_ros_b200.distributed import assemble_columns, azimuth_shard
from radarays_ros_b200.radar import RadarB200

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world():
    sc = scenes.urban_5m()
    cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, n_cells=3360, n_samples=256, n_reflections=3, include_motion=0))
    radar = RadarB200(sc, cfg, beam_seed=20240310, noise_seed=20240310)
    return sc, cfg, radar


def test_full_size_frame_equals_oracle(world, oracle_mod):
    sc, cfg, radar = world
    assert sc.n_tris >= 5_000_000
    img, st = radar.simulate(sc.pose_array()[0], frame_id=3, return_stats=True)
    o = oracle_mod.OracleScene(sc).simulate(cfg, radar.getBeamSamples(), sc.pose_array()[:1], noise_seed=20240310,
                                            frame_id=3, want_columns=False, records=True, record_capacity=400 * 256 * 8)
    assert st.n_casts == len(o["casts"]), "rays*bounces differ"
    assert st.n_hits == int((o["casts"]["face_id"] >= 0).sum())
    assert st.n_signals == len(o["signals"])
    assert np.array_equal(img, o["image"]), "full-size mono8 frame differs from the oracle"
    assert img.max() > 50


def test_full_size_properties(world):
    sc, cfg, radar = world
    poses = sc.pose_array()
    batch = radar.simulate(poses, frame_id=100)
    again = radar.simulate(poses, frame_id=100)
    assert np.array_equal(batch, again), "not deterministic"
    for i in (0, 7, 15):
        single = radar.simulate(poses[i], frame_id=100 + i)
        assert np.array_equal(batch[i], single), "batched pose %d != single call" % i
    other = radar.simulate(poses[0], frame_id=999)
    assert not np.array_equal(other, batch[0]), "noise is not keyed by the frame id"
    # azimuth shards (column-major, device-resident API) reassemble to the full frame
    dev = torch.device("cuda", 0)
    p = np.frombuffer((type(poses[0]) * 1)(poses[5]), dtype=np.float32).reshape(1, 7).copy()
    d_pose = torch.from_numpy(p).to(dev)
    shards = []
    for r in range(3):
        b, c = azimuth_shard(r, 3)
        d_cols = torch.zeros((c, cfg.n_cells), dtype=torch.uint8, device=dev)
        radar.simulate_device(d_pose.data_ptr(), 1, d_cols.data_ptr(), frame_id=105, azimuth_begin=b, azimuth_count=c,
                              column_major=True, stream=torch.cuda.current_stream(dev).cuda_stream)
        torch.cuda.synchronize()
        shards.append(d_cols.cpu().numpy())
    assert np.array_equal(assemble_columns(shards, cfg.n_cells, cfg.scroll_image), batch[5])
    st = radar.get_stats()
    assert st.overflow == 0


def test_config3_2048_samples_equals_oracle(world, oracle_mod):
    """BASELINE config 3 at the top of its sweep: urban-5M, 2 passes, 2048 beam samples per azimuth (819 200 pass-0 rays),
    one full frame against the oracle: casts, hits, returns and every pixel."""
    sc, cfg, radar = world
    cfg3 = cfg.copy().update(n_samples=2048, n_reflections=2)
    radar.updateDynCfg(cfg3)
    try:
        img, st = radar.simulate(sc.pose_array()[2], frame_id=11, return_stats=True)
        o = oracle_mod.OracleScene(sc).simulate(cfg3, radar.getBeamSamples(), sc.pose_array()[2:3], noise_seed=20240310,
                                                frame_id=11, want_columns=False, records=True, record_capacity=400 * 2048 * 3)
        assert st.n_casts == len(o["casts"]) and st.n_casts >= 400 * 2048
        assert st.n_hits == int((o["casts"]["face_id"] >= 0).sum())
        assert st.n_signals == len(o["signals"])
        assert np.array_equal(img, o["image"]), "config 3 (2048 samples) frame differs from the oracle"
    finally:
        radar.updateDynCfg(cfg)


def test_config4_warehouse_1m_equals_oracle(oracle_mod):
    """BASELINE config 4 at FULL size: warehouse-1M (~1 M triangles), 5 passes, 256 samples, dielectric / metal mix —
    the wave lists GROW here (glass and plastic wrap split every wave in two). One frame against the oracle bit for bit,
    then the same frame as 8 azimuth shards (the layout rr_simulate_sharded renders) reassembled."""
    sc = scenes.warehouse()
    assert sc.n_tris >= 900_000
    cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, n_cells=3360, n_samples=256, n_reflections=5, resolution=0.02, include_motion=0))
    radar = RadarB200(sc, cfg, beam_seed=20240310, noise_seed=20240310)
    radar.setMaxWavesPerAzimuth(256 * 10)
    pose = sc.pose_array()[3]
    img, st = radar.simulate(pose, frame_id=5, return_stats=True)
    o = oracle_mod.OracleScene(sc).simulate(cfg, radar.getBeamSamples(), sc.pose_array()[3:4], noise_seed=20240310,
                                            frame_id=5, want_columns=False, records=True, record_capacity=400 * 256 * 40)
    assert st.overflow == 0
    assert st.n_casts == len(o["casts"]), "rays*bounces differ"
    assert st.n_casts > 400 * 256 * 3, "the dielectric splits should make the lists grow"
    assert st.n_hits == int((o["casts"]["face_id"] >= 0).sum())
    assert st.n_signals == len(o["signals"])
    assert np.array_equal(img, o["image"]), "config 4 full-size frame differs from the oracle"
    dev = torch.device("cuda", 0)
    p = np.frombuffer((type(pose) * 1)(pose), dtype=np.float32).reshape(1, 7).copy()
    d_pose = torch.from_numpy(p).to(dev)
    shards = []
    for r in range(8):
        b, c = azimuth_shard(r, 8)
        d_cols = torch.zeros((c, cfg.n_cells), dtype=torch.uint8, device=dev)
        radar.simulate_device(d_pose.data_ptr(), 1, d_cols.data_ptr(), frame_id=5, azimuth_begin=b, azimuth_count=c,
                              column_major=True, stream=torch.cuda.current_stream(dev).cuda_stream)
        torch.cuda.synchronize()
        shards.append(d_cols.cpu().numpy())
    assert np.array_equal(assemble_columns(shards, cfg.n_cells, cfg.scroll_image), img)
