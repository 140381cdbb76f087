// RadarB200.hpp — the reference-side binding: a third backend next to RadarCPU / RadarGPU.
//
// Drop into uos/radarays_ros as include/radarays_ros/RadarB200.hpp; it needs only the reference's own
// Radar.hpp (ROS types) and include/radarays_b200.h + libradarays_b200.so from this repo. INTEGRATION.md shows the
// three lines of radar_simulator.cpp that select it. In this repository it is compiled against the reference's
// UNMODIFIED Radar.hpp / Radar.cpp and the ROS stand-ins of oracle/ref_shim (oracle/adapter_harness.cpp ->
// oracle/_ref/libradarays_adapter.so) and checked on the GPU by tests/test_gpu_adapter.py.
//
// Mirrors RadarCPU (include/radarays_ros/RadarCPU.hpp:17-36): same constructor shape, same
// `sensor_msgs::ImagePtr simulate(ros::Time)` contract — null pointer when TF is unavailable
// (RadarCPU.cpp:129-133), mono8 n_cells x 400 image otherwise (RadarCPU.cpp:555-561).
#ifndef RADARAYS_ROS_RADAR_B200_HPP
#define RADARAYS_ROS_RADAR_B200_HPP

#include "Radar.hpp"
#include <radarays_b200.h>
#include <sensor_msgs/image_encodings.h>
#include <stdexcept>
#include <vector>

namespace radarays_ros
{

class RadarB200 : public Radar
{
public:
    using Base = Radar;

    // verts/faces/face_objects: the triangle soup of the map (what rm::import_embree_map would load);
    // face_objects[i] = Embree geometry / instance id of face i (index into `object_materials`)
    RadarB200(std::shared_ptr<ros::NodeHandle> nh_p,
              std::shared_ptr<tf2_ros::Buffer> tf_buffer,
              std::shared_ptr<tf2_ros::TransformListener> tf_listener,
              std::string map_frame, std::string sensor_frame,
              const std::vector<float>& verts_xyz, const std::vector<uint32_t>& faces,
              const std::vector<uint32_t>& face_objects, int device = 0)
    : Base(nh_p, tf_buffer, tf_listener, map_frame, sensor_frame)
    {
        has_last = false;   // Radar.hpp:82 declares `bool has_last;` without an initialiser and Radar.cpp never sets it
                            // before updateTsm() reads it (Radar.cpp:105): without this, "TF unavailable" on the very
                            // first frame renders from the identity pose whenever the garbage happens to be non-zero
        check(rr_create(&m_ctx, device));
        check(rr_set_mesh(m_ctx, verts_xyz.data(), verts_xyz.size() / 3, faces.data(), faces.size() / 3,
                          face_objects.empty() ? nullptr : face_objects.data()));
        check(rr_set_beam_samples(m_ctx, nullptr, 0, /*seed=*/ros::Time::now().toNSec()));
    }
    ~RadarB200() { rr_destroy(m_ctx); }

    virtual sensor_msgs::ImagePtr simulate(ros::Time stamp)
    {
        sensor_msgs::ImagePtr msg;
        // include_motion == false: one TF lookup per frame (RadarCPU.cpp:127-134); null = "no frame", the caller
        // checks (radar_simulator.cpp:88).
        // include_motion == true : the reference refreshes Tsm before every azimuth (RadarCPU.cpp:190-196) and keeps
        // the previous pose when a lookup fails; here the 400 lookups happen up front and the 400 poses go to
        // rr_simulate_motion in one call.
        const bool motion = m_cfg.include_motion;
        std::vector<rr_pose> poses;
        if(!motion) {
            if(!updateTsm()) { return msg; }
            poses.push_back(to_pose(Tsm_last));
        } else {
            poses.reserve(RR_N_ANGLES);
            for(int a = 0; a < RR_N_ANGLES; a++) {
                if(!updateTsm() && !has_last) { return msg; }   // nothing to extrapolate from yet
                poses.push_back(to_pose(Tsm_last));
            }
        }

        // materials are re-read every frame by the node (radar_simulator.cpp:85,200): cheap, forward them
        std::vector<rr_material> mats(m_params.materials.data.size());
        for(size_t i = 0; i < mats.size(); i++) {
            const auto& m = m_params.materials.data[i];
            mats[i] = {m.velocity, m.ambient, m.diffuse, m.specular};
        }
        check(rr_set_materials(m_ctx, mats.data(), mats.size(), m_object_materials.data(),
                               m_object_materials.size(), m_material_id_air));

        // m_cfg (dynamic_reconfigure) -> rr_config, field by field (same names); m_params.model -> rr_model
        rr_config c; rr_config_defaults(&c);
        c.beam_width = m_cfg.beam_width; c.resolution = m_cfg.resolution; c.n_cells = m_cfg.n_cells;
        c.n_samples = m_cfg.n_samples; c.beam_sample_dist = m_cfg.beam_sample_dist;
        c.beam_sample_dist_normal_p_in_cone = m_cfg.beam_sample_dist_normal_p_in_cone;
        c.n_reflections = m_cfg.n_reflections; c.energy_max = m_cfg.energy_max; c.signal_max = m_cfg.signal_max;
        c.signal_denoising = m_cfg.signal_denoising;
        c.signal_denoising_triangular_width = m_cfg.signal_denoising_triangular_width;
        c.signal_denoising_triangular_mode = m_cfg.signal_denoising_triangular_mode;
        c.signal_denoising_gaussian_width = m_cfg.signal_denoising_gaussian_width;
        c.signal_denoising_gaussian_mode = m_cfg.signal_denoising_gaussian_mode;
        c.signal_denoising_mb_width = m_cfg.signal_denoising_mb_width;
        c.signal_denoising_mb_mode = m_cfg.signal_denoising_mb_mode;
        c.ambient_noise = m_cfg.ambient_noise;
        c.ambient_noise_at_signal_0 = m_cfg.ambient_noise_at_signal_0;
        c.ambient_noise_at_signal_1 = m_cfg.ambient_noise_at_signal_1;
        c.ambient_noise_energy_max = m_cfg.ambient_noise_energy_max;
        c.ambient_noise_energy_min = m_cfg.ambient_noise_energy_min;
        c.ambient_noise_energy_loss = m_cfg.ambient_noise_energy_loss;
        c.scroll_image = m_cfg.scroll_image; c.multipath_threshold = m_cfg.multipath_threshold;
        c.record_multi_reflection = m_cfg.record_multi_reflection; c.record_multi_path = m_cfg.record_multi_path;
        c.include_motion = motion ? 1 : 0;
        rr_model model = {(float)m_params.model.beam_width, m_params.model.n_samples, m_params.model.n_reflections};
        check(rr_set_params(m_ctx, &model, &c));             // also applies the m_resample rule (Radar.cpp:199-206)

        msg.reset(new sensor_msgs::Image());
        msg->height = m_cfg.n_cells; msg->width = RR_N_ANGLES;  // rows = range bins, cols = azimuths (Radar.cpp:34)
        msg->encoding = sensor_msgs::image_encodings::MONO8; msg->step = RR_N_ANGLES;
        msg->data.resize((size_t)msg->height * msg->width);
        if(motion) { check(rr_simulate_motion(m_ctx, poses.data(), 1, m_frame_id++, msg->data.data(), nullptr)); }
        else       { check(rr_simulate(m_ctx, poses.data(), 1, m_frame_id++, msg->data.data(), nullptr)); }
        msg->header.stamp = stamp;
        msg->header.frame_id = m_sensor_frame;
        return msg;
    }

protected:
    static rr_pose to_pose(const rm::Transform& T) { return rr_pose{T.R.x, T.R.y, T.R.z, T.R.w, T.t.x, T.t.y, T.t.z}; }
    void check(int rc) { if(rc != RR_OK) { throw std::runtime_error(rr_last_error(m_ctx)); } }
    rr_ctx* m_ctx = nullptr;
    uint64_t m_frame_id = 0;
};

using RadarB200Ptr = std::shared_ptr<RadarB200>;

} // namespace radarays_ros

#endif // RADARAYS_ROS_RADAR_B200_HPP
