// Restatement of rmagine math types (github.com/uos/rmagine, range 2.2.1...; not installed) — SURVEY.md Appendix B.
#ifndef RR_SHIM_RMAGINE_MATH_TYPES_H
#define RR_SHIM_RMAGINE_MATH_TYPES_H
#include <cmath>
namespace rmagine {
struct Quaternion; struct EulerAngles; struct Matrix3x3;

struct Vector {
    float x, y, z;
    static Vector Zeros() { return {0.0f, 0.0f, 0.0f}; }
    Vector operator+(const Vector& b) const { return {x + b.x, y + b.y, z + b.z}; }
    Vector operator-(const Vector& b) const { return {x - b.x, y - b.y, z - b.z}; }
    Vector operator-() const { return {-x, -y, -z}; }
    Vector operator*(const float& s) const { return {x * s, y * s, z * s}; }
    Vector operator/(const float& s) const { return {x / s, y / s, z / s}; }
    Vector& operator*=(const float& s) { x *= s; y *= s; z *= s; return *this; }
    float dot(const Vector& b) const { return x * b.x + y * b.y + z * b.z; }
    Vector cross(const Vector& b) const { return {y * b.z - z * b.y, z * b.x - x * b.z, x * b.y - y * b.x}; }
    float l2normSquared() const { return x * x + y * y + z * z; }
    float l2norm() const { return sqrtf(l2normSquared()); }
    Vector normalize() const { return *this / l2norm(); }
    void normalizeInplace() { const float d = l2norm(); x /= d; y /= d; z /= d; }
};
using Vector3 = Vector;
using Vector3f = Vector;
using Point = Vector;

struct Quaternion {
    float x, y, z, w;
    Quaternion inv() const { return {-x, -y, -z, w}; }
    Quaternion mult(const Quaternion& q2) const
    {
        Quaternion r;
        r.w = w * q2.w - x * q2.x - y * q2.y - z * q2.z;
        r.x = w * q2.x + x * q2.w + y * q2.z - z * q2.y;
        r.y = w * q2.y - x * q2.z + y * q2.w + z * q2.x;
        r.z = w * q2.z + x * q2.y - y * q2.x + z * q2.w;
        return r;
    }
    Vector mult(const Vector& p) const
    {
        const Quaternion P{p.x, p.y, p.z, 0.0f};
        const Quaternion PT = this->mult(P).mult(inv());
        return {PT.x, PT.y, PT.z};
    }
    Quaternion operator*(const Quaternion& q2) const { return mult(q2); }
    Vector operator*(const Vector& p) const { return mult(p); }
};

struct EulerAngles {
    float roll, pitch, yaw;
    operator Quaternion() const
    {
        const float cr = cosf(roll / 2.0f), sr = sinf(roll / 2.0f);
        const float cp = cosf(pitch / 2.0f), sp = sinf(pitch / 2.0f);
        const float cy = cosf(yaw / 2.0f), sy = sinf(yaw / 2.0f);
        Quaternion q;
        q.w = cr * cp * cy + sr * sp * sy;
        q.x = sr * cp * cy - cr * sp * sy;
        q.y = cr * sp * cy + sr * cp * sy;
        q.z = cr * cp * sy - sr * sp * cy;
        return q;
    }
    Vector operator*(const Vector& v) const { return Quaternion(*this).mult(v); }
};

struct Matrix3x3 {
    float m[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    float& operator()(int r, int c) { return m[r][c]; }
    float operator()(int r, int c) const { return m[r][c]; }
    operator Quaternion() const      // only reached from ray_to_transform (radar_algorithms.cpp:211-240), unused by the hot path
    {
        Quaternion q;
        const float tr = m[0][0] + m[1][1] + m[2][2];
        if (tr > 0) { float s = sqrtf(tr + 1.0f) * 2; q.w = 0.25f * s; q.x = (m[2][1] - m[1][2]) / s; q.y = (m[0][2] - m[2][0]) / s; q.z = (m[1][0] - m[0][1]) / s; }
        else if (m[0][0] > m[1][1] && m[0][0] > m[2][2]) { float s = sqrtf(1.0f + m[0][0] - m[1][1] - m[2][2]) * 2; q.w = (m[2][1] - m[1][2]) / s; q.x = 0.25f * s; q.y = (m[0][1] + m[1][0]) / s; q.z = (m[0][2] + m[2][0]) / s; }
        else if (m[1][1] > m[2][2]) { float s = sqrtf(1.0f + m[1][1] - m[0][0] - m[2][2]) * 2; q.w = (m[0][2] - m[2][0]) / s; q.x = (m[0][1] + m[1][0]) / s; q.y = 0.25f * s; q.z = (m[1][2] + m[2][1]) / s; }
        else { float s = sqrtf(1.0f + m[2][2] - m[0][0] - m[1][1]) * 2; q.w = (m[1][0] - m[0][1]) / s; q.x = (m[0][2] + m[2][0]) / s; q.y = (m[1][2] + m[2][1]) / s; q.z = 0.25f * s; }
        return q;
    }
};

struct Transform {
    Quaternion R;
    Vector t;
    static Transform Identity() { return {{0.0f, 0.0f, 0.0f, 1.0f}, {0.0f, 0.0f, 0.0f}}; }
    Transform inv() const { Transform r; r.R = R.inv(); r.t = -(r.R * t); return r; }
    Transform operator~() const { return inv(); }
    Transform operator*(const Transform& T2) const { Transform r; r.t = R * T2.t + t; r.R = R * T2.R; return r; }
    Vector operator*(const Vector& v) const { return R * v + t; }
};
} // namespace rmagine
#endif
