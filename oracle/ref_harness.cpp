/* ref_harness.cpp — drives the reference's OWN RadarCPU::simulate (compiled in place from /root/reference against
 * oracle/ref_shim) behind a small C interface. TEST INFRASTRUCTURE: validates oracle/rr_oracle.cpp and serves as the
 * "reference" CPU baseline of bench.py. Nothing here restates the algorithm; it only plays the role of the ROS node
 * (src/radar_simulator.cpp:83-96,145-158): construct the backend, deliver parameters, call simulate(). */
#include <radarays_ros/RadarCPU.hpp>
#include <radarays_ros/ros_helper.h>
#include <radarays_ros/radar_algorithms.h>
#include <radarays_ros/radar_math.h>
#include <omp.h>
#include <sstream>
#include "../include/radarays_b200.h"     /* POD layouts of the C interface only */

namespace rr_ref_shim {
static NoiseState g_noise;
NoiseState& noise_state() { return g_noise; }
thread_local uint32_t tls_azimuth = 0;
thread_local BeamState tls_beam;
/* standard normal draw = the reference's own quantile() (radar_math.h:47-50) of an open-interval uniform */
float std_normal_from_bits(uint32_t bits) { return radarays_ros::quantile(((float)(bits >> 9) + 0.5f) * 0x1p-23f); }
}

/* ros_helper.cpp (XmlRpc parsing) is not compiled; this is the one function of it Radar::loadParams calls */
radarays_ros::RadarMaterials loadRadarMaterialsFromParameterServer(std::shared_ptr<ros::NodeHandle> nh) { return nh->materials; }

namespace {
class Harness : public radarays_ros::RadarCPU {
public:
    Harness(std::shared_ptr<ros::NodeHandle> nh, std::shared_ptr<tf2_ros::Buffer> buf,
            std::shared_ptr<tf2_ros::TransformListener> lis, rmagine::EmbreeMapPtr map)
    : radarays_ros::RadarCPU(nh, buf, lis, "map", "navtech", map), map_(map) {}

    void deliver(const radarays_ros::RadarModelConfig& cfg) { m_dyn_rec_server.deliver(cfg); }   /* -> Radar::updateDynCfg */

    void set_beam(const float* dirs, size_t n)       /* m_waves_start as sample_cone_local builds it (RadarCPU.cpp:106-114,136-145) */
    {
        radarays_ros::DirectedWave wave;
        wave.energy = 1.0; wave.polarization = 0.5; wave.frequency = 76.5; wave.velocity = 0.3;
        wave.material_id = 0; wave.time = 0.0;
        wave.ray.orig = {0.0, 0.0, 0.0};
        m_waves_start.assign(n, wave);
        for (size_t i = 0; i < n; i++) m_waves_start[i].ray.dir = {dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]};
        m_resample = false;
    }
    void prepare_sims(int n_threads)                 /* the reference inserts into m_sims unguarded inside the OpenMP region */
    {
        for (int t = 0; t < n_threads; t++) {
            if (m_sims.find(t) == m_sims.end()) {
                m_sims[t] = std::make_shared<rmagine::OnDnSimulatorEmbree>(map_);
                m_sims[t]->setTsb(rmagine::Transform::Identity());
            }
        }
    }
    rmagine::EmbreeMapPtr map_;
};

struct RefCtx {
    orc::Scene* scene = nullptr;
    std::shared_ptr<ros::NodeHandle> nh;
    std::shared_ptr<tf2_ros::Buffer> buf;
    std::shared_ptr<tf2_ros::TransformListener> lis;
    rmagine::EmbreeMapPtr map;
    std::unique_ptr<Harness> radar;
};
}

extern "C" {

void* ref_create(const float* verts, size_t n_verts, const uint32_t* tris, size_t n_tris, const uint32_t* tri_obj)
{
    RefCtx* c = new RefCtx();
    c->scene = orc::make_scene(verts, n_verts, tris, n_tris, tri_obj);
    if (!c->scene) { delete c; return nullptr; }
    c->nh = std::make_shared<ros::NodeHandle>("~");
    c->buf = std::make_shared<tf2_ros::Buffer>();
    c->lis = std::make_shared<tf2_ros::TransformListener>(*c->buf);
    c->map = std::make_shared<rmagine::EmbreeMap>();
    c->map->scene = c->scene;
    return c;
}

void ref_destroy(void* h)
{
    RefCtx* c = (RefCtx*)h;
    if (!c) return;
    c->radar.reset();
    delete c->scene;
    delete c;
}

/* poses: 1 (include_motion = false) or 400 (include_motion = true). Returns 0, or 1 when simulate() returned null. */
int ref_simulate(void* h, const rr_config* cfg, const rr_model* model,
                 const rr_material* materials, size_t n_materials, const int32_t* object_materials, size_t n_objects,
                 int32_t material_id_air, const float* beam_dirs, size_t n_dirs,
                 const rr_pose* poses, size_t n_poses, uint64_t noise_seed, uint64_t frame_id,
                 int n_threads, int brute_force, uint8_t* out_polar, double* elapsed_s)
{
    RefCtx* c = (RefCtx*)h;
    /* parameter server content (Radar::loadParams, Radar.cpp:220-226) */
    c->nh->materials.data.resize(n_materials);
    for (size_t i = 0; i < n_materials; i++) {
        c->nh->materials.data[i].velocity = materials[i].velocity; c->nh->materials.data[i].ambient = materials[i].ambient;
        c->nh->materials.data[i].diffuse = materials[i].diffuse; c->nh->materials.data[i].specular = materials[i].specular;
    }
    c->nh->object_materials.assign(object_materials, object_materials + n_objects);
    c->nh->material_id_air = material_id_air;
    c->map->brute_force = brute_force != 0;
    if (!c->radar) c->radar.reset(new Harness(c->nh, c->buf, c->lis, c->map));
    c->radar->loadParams();

    /* dynamic_reconfigure request (Radar::updateDynCfg, Radar.cpp:188-218) */
    radarays_ros::RadarModelConfig g;
    g.z_offset = cfg->z_offset; g.range_min = cfg->range_min; g.range_max = cfg->range_max; g.beam_width = cfg->beam_width;
    g.resolution = cfg->resolution; g.n_cells = cfg->n_cells; g.n_samples = cfg->n_samples; g.beam_sample_dist = cfg->beam_sample_dist;
    g.beam_sample_dist_normal_p_in_cone = cfg->beam_sample_dist_normal_p_in_cone; g.n_reflections = cfg->n_reflections;
    g.energy_min = cfg->energy_min; g.energy_max = cfg->energy_max; g.signal_max = cfg->signal_max;
    g.signal_denoising = cfg->signal_denoising;
    g.signal_denoising_triangular_width = cfg->signal_denoising_triangular_width; g.signal_denoising_triangular_mode = cfg->signal_denoising_triangular_mode;
    g.signal_denoising_gaussian_width = cfg->signal_denoising_gaussian_width; g.signal_denoising_gaussian_mode = cfg->signal_denoising_gaussian_mode;
    g.signal_denoising_mb_width = cfg->signal_denoising_mb_width; g.signal_denoising_mb_mode = cfg->signal_denoising_mb_mode;
    g.ambient_noise = cfg->ambient_noise; g.ambient_noise_at_signal_0 = cfg->ambient_noise_at_signal_0;
    g.ambient_noise_at_signal_1 = cfg->ambient_noise_at_signal_1; g.ambient_noise_energy_max = cfg->ambient_noise_energy_max;
    g.ambient_noise_energy_min = cfg->ambient_noise_energy_min; g.ambient_noise_energy_loss = cfg->ambient_noise_energy_loss;
    g.ambient_noise_uniform_max = cfg->ambient_noise_uniform_max; g.ambient_noise_perlin_scale_low = cfg->ambient_noise_perlin_scale_low;
    g.ambient_noise_perlin_scale_high = cfg->ambient_noise_perlin_scale_high; g.ambient_noise_perlin_p_low = cfg->ambient_noise_perlin_p_low;
    g.scroll_image = cfg->scroll_image; g.multipath_threshold = cfg->multipath_threshold;
    g.record_multi_reflection = cfg->record_multi_reflection != 0; g.record_multi_path = cfg->record_multi_path != 0;
    g.include_motion = cfg->include_motion != 0;
    c->radar->deliver(g);
    if (model) {                                      /* Radar::setParams (Radar.hpp:56-59) */
        radarays_ros::RadarParams p = c->radar->getParams();
        p.model.beam_width = model->beam_width; p.model.n_samples = model->n_samples; p.model.n_reflections = model->n_reflections;
        c->radar->setParams(p);
    }
    c->radar->set_beam(beam_dirs, n_dirs);

    /* tf: one transform per lookup */
    c->buf->queue.clear(); c->buf->next = 0;
    for (size_t i = 0; i < n_poses; i++) {
        geometry_msgs::TransformStamped t;
        t.transform.translation.x = poses[i].tx; t.transform.translation.y = poses[i].ty; t.transform.translation.z = poses[i].tz;
        t.transform.rotation.x = poses[i].qx; t.transform.rotation.y = poses[i].qy; t.transform.rotation.z = poses[i].qz; t.transform.rotation.w = poses[i].qw;
        c->buf->queue.push_back(t);
    }
    rr_ref_shim::g_noise.seed = noise_seed; rr_ref_shim::g_noise.frame = frame_id;
    if (n_threads > 0) omp_set_num_threads(n_threads);
    c->radar->prepare_sims(omp_get_max_threads());

    /* the reference reports its own timing on stdout (RadarCPU.cpp:147-148,550-553): capture it */
    std::ostringstream captured;
    std::streambuf* old = std::cout.rdbuf(captured.rdbuf());
    sensor_msgs::ImagePtr msg = c->radar->simulate(ros::Time(0.0));
    std::cout.rdbuf(old);
    if (elapsed_s) {
        double last = 0.0; std::string line; std::istringstream is(captured.str());
        while (std::getline(is, line)) { char* end = nullptr; const double v = std::strtod(line.c_str(), &end); if (end != line.c_str() && *end == '\0') last = v; }
        *elapsed_s = last;
    }
    if (!msg) return 1;
    if (out_polar) std::memcpy(out_polar, msg->data.data(), msg->data.size());
    return 0;
}

/* The reference's OWN sample_cone_local (radar_algorithms.cpp:248-294), its std::mt19937 / distributions replaced by the
 * Philox-fed shims of pre.h: pins oracle::sample_cone_local and the library's draw_beam_samples (as SETS of directions:
 * both store the i.i.d. draws along a Morton curve, the reference in draw order). */
int ref_sample_cone_local(float width, int n_samples, int sample_dist, float p_in_cone, uint64_t seed, float* dirs_out)
{
    radarays_ros::DirectedWave wave;
    wave.energy = 1.0; wave.polarization = 0.5; wave.frequency = 76.5; wave.velocity = 0.3;
    wave.material_id = 0; wave.time = 0.0;
    wave.ray.orig = {0.0, 0.0, 0.0}; wave.ray.dir = {1.0, 0.0, 0.0};
    rr_ref_shim::tls_beam.active = true; rr_ref_shim::tls_beam.seed = seed;
    const std::vector<radarays_ros::DirectedWave> waves = radarays_ros::sample_cone_local(wave, width, n_samples, sample_dist, p_in_cone);
    rr_ref_shim::tls_beam.active = false;
    for (size_t i = 0; i < waves.size(); i++) {
        dirs_out[3 * i] = waves[i].ray.dir.x; dirs_out[3 * i + 1] = waves[i].ray.dir.y; dirs_out[3 * i + 2] = waves[i].ray.dir.z;
    }
    return (int)waves.size();
}

} // extern "C"
