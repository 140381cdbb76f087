/* rr_bvh_build.cu — BVH construction entry point used by rr_set_mesh. */
#include <chrono>
#include <string>
#include "rr_bvh.h"

int rr_bvh_build_device(const RRTriSoup& soup, RRPackedBVH& out, float* build_ms, std::string& err)
{
    (void)err;
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<RRBuildNode> nodes;
    std::vector<uint32_t> order;
    rr_bvh_build_host(soup, nodes, order);
    rr_bvh_pack(soup, nodes, order, out);
    if (build_ms) *build_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return RR_OK;
}
