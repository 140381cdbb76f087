#ifndef RR_SHIM_RMAGINE_STOPWATCH_HPP
#define RR_SHIM_RMAGINE_STOPWATCH_HPP
#include <chrono>
namespace rmagine {
class StopWatch {
public:
    double operator()()
    {
        const auto now = std::chrono::steady_clock::now();
        const double s = std::chrono::duration<double>(now - m_last).count();
        m_last = now;
        return s;
    }
private:
    std::chrono::steady_clock::time_point m_last = std::chrono::steady_clock::now();
};
}
#endif
