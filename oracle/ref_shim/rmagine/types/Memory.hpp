#ifndef RR_SHIM_RMAGINE_MEMORY_HPP
#define RR_SHIM_RMAGINE_MEMORY_HPP
#include <cstddef>
#include <vector>
namespace rmagine {
struct RAM {};
template <typename T, typename MemT = RAM>
class MemView {
public:
    MemView() = default;
    MemView(T* p, size_t n) : m_p(p), m_n(n) {}
    T& operator[](size_t i) { return m_p[i]; }
    const T& operator[](size_t i) const { return m_p[i]; }
    size_t size() const { return m_n; }
    T* raw() { return m_p; }
    const T* raw() const { return m_p; }
protected:
    T* m_p = nullptr; size_t m_n = 0;
};
template <typename T, typename MemT = RAM>
class Memory : public MemView<T, MemT> {
public:
    Memory() = default;
    explicit Memory(size_t n) { resize(n); }
    Memory(const Memory& o) : MemView<T, MemT>(), m_v(o.m_v) { sync(); }
    Memory& operator=(const Memory& o) { m_v = o.m_v; sync(); return *this; }
    void resize(size_t n) { m_v.resize(n); sync(); }
private:
    void sync() { this->m_p = m_v.data(); this->m_n = m_v.size(); }
    std::vector<T> m_v;
};
} // namespace rmagine
#endif
