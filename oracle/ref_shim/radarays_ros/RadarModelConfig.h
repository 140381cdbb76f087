#pragma once
// What dynamic_reconfigure generates from cfg/RadarModel.cfg:11-85 (field names and C++ types only).
namespace radarays_ros {
struct RadarModelConfig {
    double z_offset = 0, range_min = 0, range_max = 0, beam_width = 0, resolution = 0;
    int n_cells = 0, n_samples = 0, beam_sample_dist = 0;
    double beam_sample_dist_normal_p_in_cone = 0;
    int n_reflections = 0;
    double energy_min = 0, energy_max = 0, signal_max = 0;
    int signal_denoising = 0;
    int signal_denoising_triangular_width = 0; double signal_denoising_triangular_mode = 0;
    int signal_denoising_gaussian_width = 0; double signal_denoising_gaussian_mode = 0;
    int signal_denoising_mb_width = 0; double signal_denoising_mb_mode = 0;
    int ambient_noise = 0;
    double ambient_noise_at_signal_0 = 0, ambient_noise_at_signal_1 = 0, ambient_noise_energy_max = 0,
           ambient_noise_energy_min = 0, ambient_noise_energy_loss = 0, ambient_noise_uniform_max = 0,
           ambient_noise_perlin_scale_low = 0, ambient_noise_perlin_scale_high = 0, ambient_noise_perlin_p_low = 0;
    int scroll_image = 0;
    double multipath_threshold = 0;
    bool record_multi_reflection = false, record_multi_path = false, include_motion = false;
};
}
