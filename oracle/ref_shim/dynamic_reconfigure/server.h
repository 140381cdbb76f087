#pragma once
// dynamic_reconfigure::Server stand-in: stores the callback Radar's ctor registers (Radar.cpp:38-40) so the harness
// can deliver a RadarModelConfig exactly the way a reconfigure request would (-> Radar::updateDynCfg).
#include <cstdint>
#include <functional>
namespace boost { using std::bind; }
using namespace std::placeholders;
namespace dynamic_reconfigure {
template <typename ConfigT>
class Server {
public:
    using CallbackType = std::function<void(ConfigT&, uint32_t)>;
    void setCallback(const CallbackType& f) { m_cb = f; }
    void deliver(ConfigT cfg, uint32_t level = 0) { if (m_cb) m_cb(cfg, level); }
private:
    CallbackType m_cb;
};
}
