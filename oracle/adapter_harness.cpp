/* adapter_harness.cpp — compiles the reference-side binding radarays_ros_b200/cpp/RadarB200.hpp against the reference's
 * UNMODIFIED Radar.hpp / Radar.cpp (in place, /root/reference) and the ROS stand-ins of oracle/ref_shim, and plays the
 * role of the ROS node (src/radar_simulator.cpp:83-96,145-176): construct the backend, deliver parameters, publish TF,
 * call simulate(). TEST INFRASTRUCTURE: proves that the adapter a maintainer would add builds against the reference's
 * own base class and renders the same image as the Python mirror (tests/test_gpu_adapter.py). */
#include "../radarays_ros_b200/cpp/RadarB200.hpp"   /* finds the reference's Radar.hpp through -I$REF/include/radarays_ros */
#include <radarays_ros/ros_helper.h>
#include <cstring>

namespace rr_ref_shim {
static NoiseState g_noise;
NoiseState& noise_state() { return g_noise; }
thread_local uint32_t tls_azimuth = 0;
}

radarays_ros::RadarMaterials loadRadarMaterialsFromParameterServer(std::shared_ptr<ros::NodeHandle> nh) { return nh->materials; }

namespace {
class Harness : public radarays_ros::RadarB200 {
public:
    using radarays_ros::RadarB200::RadarB200;
    void deliver(const radarays_ros::RadarModelConfig& cfg) { m_dyn_rec_server.deliver(cfg); }   /* -> Radar::updateDynCfg */
    void seeds(uint64_t beam_seed, uint64_t noise_seed, uint64_t frame_id)
    {
        check(rr_set_beam_samples(m_ctx, nullptr, 0, beam_seed));
        check(rr_set_noise_seed(m_ctx, noise_seed));
        m_frame_id = frame_id;
    }
};

struct AdapterCtx {
    std::shared_ptr<ros::NodeHandle> nh;
    std::shared_ptr<tf2_ros::Buffer> buf;
    std::shared_ptr<tf2_ros::TransformListener> lis;
    std::vector<float> verts; std::vector<uint32_t> faces, objs;
    std::unique_ptr<Harness> radar;
    std::string err;
};
}

extern "C" {

void* adapter_create(const float* verts, size_t n_verts, const uint32_t* tris, size_t n_tris, const uint32_t* tri_obj)
{
    AdapterCtx* c = new AdapterCtx();
    c->nh = std::make_shared<ros::NodeHandle>("~");
    c->buf = std::make_shared<tf2_ros::Buffer>();
    c->lis = std::make_shared<tf2_ros::TransformListener>(*c->buf);
    c->verts.assign(verts, verts + 3 * n_verts);
    c->faces.assign(tris, tris + 3 * n_tris);
    if (tri_obj) c->objs.assign(tri_obj, tri_obj + n_tris);
    return c;
}

void adapter_destroy(void* h) { delete (AdapterCtx*)h; }

const char* adapter_last_error(void* h) { return ((AdapterCtx*)h)->err.c_str(); }

/* n_poses TF answers are queued (1, or 400 with include_motion); n_poses == 0 = "TF unavailable".
 * Returns 0 = image written, 1 = simulate() returned null (no frame), -1 = exception (adapter_last_error). */
int adapter_simulate(void* h, const rr_config* cfg, const rr_model* model,
                     const rr_material* materials, size_t n_materials, const int32_t* object_materials, size_t n_objects,
                     int32_t material_id_air, const rr_pose* poses, size_t n_poses,
                     uint64_t beam_seed, uint64_t noise_seed, uint64_t frame_id, uint8_t* out_polar)
{
    AdapterCtx* c = (AdapterCtx*)h;
    try {
        c->nh->materials.data.resize(n_materials);
        for (size_t i = 0; i < n_materials; i++) {
            c->nh->materials.data[i].velocity = materials[i].velocity; c->nh->materials.data[i].ambient = materials[i].ambient;
            c->nh->materials.data[i].diffuse = materials[i].diffuse; c->nh->materials.data[i].specular = materials[i].specular;
        }
        c->nh->object_materials.assign(object_materials, object_materials + n_objects);
        c->nh->material_id_air = material_id_air;
        if (!c->radar) c->radar.reset(new Harness(c->nh, c->buf, c->lis, "map", "navtech", c->verts, c->faces, c->objs, 0));
        c->radar->loadParams();                               /* the node does this before every frame, radar_simulator.cpp:85 */

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
        if (model) {                                          /* Radar::setParams (Radar.hpp:56-59) */
            radarays_ros::RadarParams p = c->radar->getParams();
            p.model.beam_width = model->beam_width; p.model.n_samples = model->n_samples; p.model.n_reflections = model->n_reflections;
            c->radar->setParams(p);
        }
        c->radar->seeds(beam_seed, noise_seed, frame_id);

        c->buf->queue.clear(); c->buf->next = 0;
        for (size_t i = 0; i < n_poses; i++) {
            geometry_msgs::TransformStamped t;
            t.transform.translation.x = poses[i].tx; t.transform.translation.y = poses[i].ty; t.transform.translation.z = poses[i].tz;
            t.transform.rotation.x = poses[i].qx; t.transform.rotation.y = poses[i].qy; t.transform.rotation.z = poses[i].qz; t.transform.rotation.w = poses[i].qw;
            c->buf->queue.push_back(t);
        }
        sensor_msgs::ImagePtr msg = c->radar->simulate(ros::Time(0.0));
        if (!msg) return 1;
        if (msg->encoding != "mono8" || msg->width != (uint32_t)RR_N_ANGLES || msg->height != (uint32_t)cfg->n_cells
            || msg->step != (uint32_t)RR_N_ANGLES || msg->header.frame_id != "navtech") { c->err = "unexpected sensor_msgs::Image layout"; return -1; }
        if (out_polar) std::memcpy(out_polar, msg->data.data(), msg->data.size());
        return 0;
    } catch (const std::exception& e) {
        c->err = e.what();
        return -1;
    }
}

} // extern "C"
