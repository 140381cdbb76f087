"""GPU edge cases of the launch-wide wavefront (rr_internal.h): empty lists, ragged list lengths, list overflow,
sub-batch / lane splitting. All comparisons are bit-exact against the CPU oracle or against the single-pose result."""
import numpy as np
import pytest

from radarays_ros_b200 import RadarModelConfig, MULRAN_DYNCFG, Pose, scenes
from radarays_ros_b200.capi import RadaRaysError
from radarays_ros_b200.radar import RadarB200
from radarays_ros_b200.scenes import Scene

pytestmark = pytest.mark.gpu


def _one_triangle_scene(z):
    """A single small triangle far below the sensor plane: every beam ray misses."""
    v = np.array([[0, 0, z], [1, 0, z], [0, 1, z]], np.float32)
    t = np.array([[0, 1, 2]], np.uint32)
    return Scene("one-triangle", v, t, np.zeros(1, np.uint32), [(0.3, 1.0, 0.0, 1.0), (0.0, 1.0, 0.0, 3000.0)], [1],
                 0, [(0.0, 0.0, 0.0, 0.0)])


@pytest.mark.parametrize("noise", [0, 2])
def test_all_rays_miss(oracle_mod, noise):
    """pass 0 produces no child: the lists of passes 1.. are empty (pass_total = 0) and every column is noise only
    (quirk 14: max_val = 0 -> 0/0 -> NaN -> 0)."""
    sc = _one_triangle_scene(-500.0)
    cfg = RadarModelConfig(n_reflections=3, n_samples=40, ambient_noise=noise, include_motion=0, n_cells=512)
    radar = RadarB200(sc, cfg, beam_seed=5, noise_seed=6)
    img, st = radar.simulate(sc.pose_array()[0], frame_id=1, return_stats=True)
    assert st.n_casts == 400 * 40 and st.n_hits == 0 and st.n_signals == 0
    o = oracle_mod.OracleScene(sc).simulate(cfg, radar.getBeamSamples(), sc.pose_array()[:1], noise_seed=6, frame_id=1)
    assert np.array_equal(img, o["image"])


def test_ragged_lists_and_lane_split(oracle_mod):
    """n_samples = 37 (not a multiple of the 32-wave group), dielectric splits (warehouse glass) so that azimuth runs
    straddle groups; 5 poses so that the sub-batches of the two lanes are uneven. One lane == two lanes == per pose."""
    sc = scenes.warehouse_small()
    cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, n_samples=37, n_reflections=4, n_cells=900, resolution=0.03))
    radar = RadarB200(sc, cfg, beam_seed=3, noise_seed=4)
    radar.setMaxWavesPerAzimuth(37 * 16)
    poses = sc.pose_array(5)
    two = radar.simulate(poses, frame_id=50).copy()
    radar.setLanes(1)
    one = radar.simulate(poses, frame_id=50).copy()
    radar.setLanes(2)
    assert np.array_equal(one, two), "result depends on the number of lanes"
    osc = oracle_mod.OracleScene(sc)
    dirs = radar.getBeamSamples()
    for i in (0, 4):
        o = osc.simulate(cfg, dirs, poses[i:i + 1], noise_seed=4, frame_id=50 + i)
        assert np.array_equal(two[i], o["image"]), "pose %d of the batch differs from the oracle" % i
        single = radar.simulate(poses[i], frame_id=50 + i)
        assert np.array_equal(single, two[i])


def test_wave_list_overflow_is_reported():
    """A list longer than max_waves_per_azimuth * azimuths must come back as RR_ERR_WAVE_OVERFLOW (-5), not as a
    silently truncated image (the reference's std::vector just grows, RadarCPU.cpp:380-389)."""
    sc = scenes.warehouse_small()
    cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, n_samples=64, n_reflections=5, n_cells=900, resolution=0.03))
    radar = RadarB200(sc, cfg, beam_seed=3, noise_seed=4)
    _, st = radar.simulate(sc.pose_array()[0], frame_id=0, return_stats=True)
    assert st.overflow == 0
    assert st.max_waves >= 64
    # capacity below what pass 0 itself needs is clamped to n_samples; make later passes overflow instead
    if st.n_casts > 400 * 64 * 5:        # the scene splits waves somewhere: total casts exceed samples * passes
        radar.setMaxWavesPerAzimuth(64)
        with pytest.raises(RadaRaysError) as ei:
            radar.simulate(sc.pose_array()[0], frame_id=0)
        assert ei.value.code == -5
        radar.setMaxWavesPerAzimuth(64 * 16)
        img = radar.simulate(sc.pose_array()[0], frame_id=0)
        assert img.max() > 0


def test_caller_buffer_and_pinned_output(oracle_mod):
    """rr_simulate writes into a caller-owned buffer; a page-locked one is filled by direct device->host copies."""
    import torch
    sc = scenes.urban_small()
    cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, n_samples=16, n_reflections=2, n_cells=768))
    radar = RadarB200(sc, cfg, beam_seed=1, noise_seed=2)
    poses = sc.pose_array(6)
    ref = radar.simulate(poses, frame_id=10).copy()
    pinned = torch.zeros((6, 768, 400), dtype=torch.uint8).pin_memory()
    out = radar.simulate(poses, frame_id=10, out=pinned.numpy())
    assert np.array_equal(out, ref) and np.array_equal(pinned.numpy(), ref)
    pageable = np.zeros((6, 768, 400), np.uint8)
    radar.simulate(poses, frame_id=10, out=pageable)
    assert np.array_equal(pageable, ref)
    with pytest.raises(ValueError):
        radar.simulate(poses, frame_id=10, out=np.zeros((5, 768, 400), np.uint8))


def test_map_file_equals_arrays(tmp_path):
    """setMapFile (rr_set_mesh_file: .ply reader + upload + BVH build) renders the same image as the arrays."""
    sc = scenes.box_room_cylinder()
    path = tmp_path / "room.ply"
    scenes.write_ply(path, sc.verts, sc.tris)
    cfg = RadarModelConfig(n_reflections=2, ambient_noise=2, include_motion=0)
    a = RadarB200(None, cfg, beam_seed=1, noise_seed=2)
    assert a.setMapFile(path) == 1
    a.loadParams(sc.materials, [1], sc.material_id_air)
    b = RadarB200(None, cfg, beam_seed=1, noise_seed=2)
    b.setMap(sc.verts, sc.tris, np.zeros(sc.n_tris, np.uint32))
    b.loadParams(sc.materials, [1], sc.material_id_air)
    pose = sc.pose_array()[0]
    img = a.simulate(pose, frame_id=3)
    assert img.max() > 0 and np.array_equal(img, b.simulate(pose, frame_id=3))


def test_cpp_example_program(tmp_path):
    """radarays_ros_b200/cpp/example_render.cpp — the C ABI from plain C++ (map file in, PGM out) — renders the image the
    Python mirror renders for the same map, parameters and pose."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "radarays_ros_b200", "example_render")
    if not os.path.exists(exe):
        pytest.skip("example_render not built (make -C radarays_ros_b200/csrc)")
    sc = scenes.box_room_cylinder()
    ply, pgm = tmp_path / "room.ply", tmp_path / "out.pgm"
    scenes.write_ply(ply, sc.verts, sc.tris)
    x, y, z, yaw = 1.5, -2.0, 1.0, 0.3
    out = subprocess.run([exe, str(ply), str(pgm), str(x), str(y), str(z), str(yaw), "3"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "3 frame(s)" in out.stdout
    data = pgm.read_bytes()
    cfg = RadarModelConfig(include_motion=0)
    header = b"P5\n400 %d\n255\n" % cfg.n_cells
    assert data.startswith(header)
    img = np.frombuffer(data[len(header):], np.uint8).reshape(cfg.n_cells, 400)
    radar = RadarB200(None, cfg, beam_seed=1, noise_seed=0)
    radar.setMapFile(ply)
    radar.loadParams([(0.3, 1.0, 0.0, 1.0), (0.0, 1.0, 0.0, 3000.0)], [1], 0)
    want = radar.simulate(Pose.from_xyz_yaw(x, y, z, yaw), frame_id=0)
    assert np.array_equal(img, want) and img.max() > 0


@pytest.mark.parametrize("n_cells", [257, 1001, 3360, 10000])
def test_row_major_output_paths_equal_the_oracle(oracle_mod, n_cells):
    """The image leaves the draw kernel either as 8-byte row segments (groups of 8 adjacent azimuths, the last CTA of a
    group transposes the staged columns; needs scroll % 8 == 0) or as single bytes (any other scroll): both against the
    oracle for column lengths that are not multiples of 4 / 16 and for the 10000-cell maximum, on a pose batch."""
    sc = scenes.box_room_cylinder()
    for scroll in (0, 8, 392, 3):
        cfg = RadarModelConfig(n_reflections=2, include_motion=0, n_cells=n_cells, resolution=30.0 / n_cells, n_samples=12,
                               ambient_noise=2, scroll_image=scroll)
        radar = RadarB200(sc, cfg, beam_seed=4, noise_seed=5)
        poses = sc.pose_array(3)
        imgs = radar.simulate(poses, frame_id=60)
        osc = oracle_mod.OracleScene(sc)
        for i in range(3):
            o = osc.simulate(cfg, radar.getBeamSamples(), poses[i:i + 1], noise_seed=5, frame_id=60 + i)
            assert np.array_equal(imgs[i], o["image"]), "n_cells %d scroll %d pose %d" % (n_cells, scroll, i)


@pytest.mark.parametrize("switch", ["RR_PASS_SPLIT", "RR_PASS_DUAL"])
def test_alternative_pass_kernels_match_the_fused_kernel(oracle_mod, monkeypatch, switch):
    """RR_PASS_SPLIT=1: every pass runs as rr_walk_kernel (cast only, hit records) + rr_shade_kernel; RR_PASS_DUAL=1: as
    rr_dual_kernel (two rays per lane, both walks advanced by one branch-free step). Both are measured alternatives to
    the fused rr_trace_kernel read when a context first sizes its scratch. Same images, same counters."""
    sc = scenes.warehouse_small()
    cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, n_samples=37, n_reflections=4, n_cells=900, resolution=0.03,
                                  record_multi_path=1, record_multi_reflection=1))
    poses = sc.pose_array(3)
    out = {}
    monkeypatch.delenv("RR_PASS_SPLIT", raising=False)
    monkeypatch.delenv("RR_PASS_DUAL", raising=False)
    for on in ("0", "1"):
        monkeypatch.setenv(switch, on)
        radar = RadarB200(sc, cfg, beam_seed=3, noise_seed=4)
        radar.setMaxWavesPerAzimuth(37 * 16)
        img = radar.simulate(poses, frame_id=9).copy()
        one, st = radar.simulate_stats(poses[1], frame_id=10)
        out[on] = (img, one, (st.n_casts, st.n_hits, st.n_signals, st.nodes_visited, st.tris_tested), radar.kernel_launches())
    assert np.array_equal(out["0"][0], out["1"][0]) and np.array_equal(out["0"][1], out["1"][1])
    assert out["0"][2] == out["1"][2]
    if switch == "RR_PASS_SPLIT":
        assert out["1"][3] > out["0"][3], "the split form launches one more kernel per pass"
    o = oracle_mod.OracleScene(sc).simulate(cfg, radar.getBeamSamples(), poses[1:2], noise_seed=4, frame_id=10)
    assert np.array_equal(out["1"][1], o["image"])


@pytest.mark.parametrize("pinned", [False, True])
def test_host_call_sub_batches_and_copy_stream(pinned):
    """A 70-pose host call is cut into sub-batches 18 18 18 8 8 (rr_api.cu: equal cuts, the last one halved down to 8 so
    that the only exposed image copy is small) whose images leave on the copy stream while the next sub-batch computes.
    Every frame must equal the same pose rendered alone with the same frame id, for a page-locked and a pageable buffer."""
    import torch
    sc = scenes.box_room_cylinder()
    cfg = RadarModelConfig(n_reflections=2, n_samples=8, n_cells=256, ambient_noise=2, include_motion=0)
    radar = RadarB200(sc, cfg, beam_seed=11, noise_seed=12)
    base = sc.pose_array()[0]
    poses = (Pose * 70)()
    for i in range(70):
        poses[i] = Pose.from_xyz_yaw(base.tx + 0.05 * i, base.ty - 0.03 * i, base.tz, 0.01 * i)
    out = torch.empty((70, 256, 400), dtype=torch.uint8, pin_memory=True).numpy() if pinned else np.empty((70, 256, 400), np.uint8)
    out[:] = 7
    imgs = radar.simulate(poses, frame_id=1000, out=out)
    assert imgs.shape == (70, 256, 400)
    for i in (0, 17, 18, 35, 36, 53, 54, 61, 62, 69):
        single = radar.simulate(poses[i], frame_id=1000 + i)
        assert np.array_equal(single, imgs[i]), "frame %d of the 70-pose call differs from the pose rendered alone" % i
    assert len({imgs[i].tobytes() for i in (0, 18, 54, 62, 69)}) == 5
