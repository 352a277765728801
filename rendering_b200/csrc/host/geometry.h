// geometry.h — float vector / matrix value types of the host library.
//
// Same public names as the reference's include/geometry.h (Vec2f, Vec3f, Matrix44f, Ray with
// dotProduct / crossProduct / length / normalize / multVecMatrix) so scene-graph code reads the
// same, but written as plain non-template float structs.  Every operation evaluates in the same
// order and precision as the reference so loaders and the tree builder produce bit-identical
// inputs for the GPU path:
//   * sums of products associate left to right, no FMA contraction (build with -ffp-contract=off);
//   * normalize() scales by (float)(1.0 / sqrt((double)len2)) — the reference's unqualified
//     `sqrt` binds to the double overload (include/geometry.h:99-112);
//   * multVecMatrix() is row-vector x matrix, adds row 3, divides by w only when w is not 0 or 1
//     (include/geometry.h:289-307).
#pragma once

#include <cmath>
#include <cstdint>

struct Vec2f {
    float x = 0.0f, y = 0.0f;
    Vec2f() = default;
    Vec2f(float s) : x(s), y(s) {}
    Vec2f(float ax, float ay) : x(ax), y(ay) {}
    Vec2f operator+(const Vec2f& o) const { return { x + o.x, y + o.y }; }
    Vec2f operator-(const Vec2f& o) const { return { x - o.x, y - o.y }; }
    Vec2f operator*(float s) const { return { x * s, y * s }; }
    Vec2f operator/(float s) const { return { x / s, y / s }; }
    float& operator[](int i) { return i == 0 ? x : y; }
    float operator[](int i) const { return i == 0 ? x : y; }
};
inline Vec2f operator*(float s, const Vec2f& v) { return { v.x * s, v.y * s }; }

struct Vec3f {
    float x = 0.0f, y = 0.0f, z = 0.0f;
    Vec3f() = default;
    Vec3f(float s) : x(s), y(s), z(s) {}
    Vec3f(float ax, float ay, float az) : x(ax), y(ay), z(az) {}

    float dotProduct(const Vec3f& o) const { float s = x * o.x; s = s + y * o.y; s = s + z * o.z; return s; }
    Vec3f crossProduct(const Vec3f& o) const
    {
        return { y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x };
    }
    float length2() const { float s = x * x; s = s + y * y; s = s + z * z; return s; }
    float length() const { return (float)std::sqrt((double)length2()); }
    Vec3f& normalize()
    {
        const float l2 = length2();
        if (l2 > 0) {
            const float k = (float)(1.0 / std::sqrt((double)l2));
            x *= k; y *= k; z *= k;
        }
        return *this;
    }
    Vec3f operator-() const { return { -x, -y, -z }; }
    Vec3f operator+(const Vec3f& o) const { return { x + o.x, y + o.y, z + o.z }; }
    Vec3f operator-(const Vec3f& o) const { return { x - o.x, y - o.y, z - o.z }; }
    Vec3f operator*(const Vec3f& o) const { return { x * o.x, y * o.y, z * o.z }; }
    Vec3f operator/(const Vec3f& o) const { return { x / o.x, y / o.y, z / o.z }; }
    Vec3f operator*(float s) const { return { x * s, y * s, z * s }; }
    Vec3f operator/(float s) const { return { x / s, y / s, z / s }; }
    Vec3f& operator+=(const Vec3f& o) { x += o.x; y += o.y; z += o.z; return *this; }
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3f operator*(float s, const Vec3f& v) { return { v.x * s, v.y * s, v.z * s }; }

struct Matrix44f {
    float m[4][4] = { { 1, 0, 0, 0 }, { 0, 1, 0, 0 }, { 0, 0, 1, 0 }, { 0, 0, 0, 1 } };

    Matrix44f() = default;
    const float* operator[](int r) const { return m[r]; }
    float* operator[](int r) { return m[r]; }

    Matrix44f operator*(const Matrix44f& b) const
    {
        Matrix44f c;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                float s = m[i][0] * b.m[0][j];
                s = s + m[i][1] * b.m[1][j];
                s = s + m[i][2] * b.m[2][j];
                s = s + m[i][3] * b.m[3][j];
                c.m[i][j] = s;
            }
        return c;
    }
    Matrix44f transposed() const
    {
        Matrix44f t;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) t.m[i][j] = m[j][i];
        return t;
    }
    Vec3f multVecMatrix(const Vec3f& s) const
    {
        float o[4];
        for (int j = 0; j < 4; ++j) {
            float a = s.x * m[0][j];
            a = a + s.y * m[1][j];
            a = a + s.z * m[2][j];
            a = a + m[3][j];
            o[j] = a;
        }
        Vec3f d(o[0], o[1], o[2]);
        const float w = o[3];
        if (w != 0.0f && w != 1.0f) {
            const float wi = 1.0f / w;
            d.x *= wi; d.y *= wi; d.z *= wi;
        }
        return d;
    }
    // Euler rotation in degrees, composed mz*my*mx exactly like Camera::getRay (scene.cpp:24-48)
    // and Mesh::loadOBJ (objects.cpp:180-204): degToRad is f*(float)M_PI/180.0f, host sinf/cosf.
    static Matrix44f rotationDeg(const Vec3f& deg)
    {
        const float pi = (float)(3.14159265358979323846);
        const float ax = deg.x * pi / 180.0f, ay = deg.y * pi / 180.0f, az = deg.z * pi / 180.0f;
        Matrix44f rx, ry, rz;
        rx.m[1][1] = cosf(ax); rx.m[1][2] = -sinf(ax); rx.m[2][1] = sinf(ax); rx.m[2][2] = cosf(ax);
        ry.m[0][0] = cosf(ay); ry.m[0][2] = sinf(ay); ry.m[2][0] = -sinf(ay); ry.m[2][2] = cosf(ay);
        rz.m[0][0] = cosf(az); rz.m[0][1] = -sinf(az); rz.m[1][0] = sinf(az); rz.m[1][1] = cosf(az);
        return (rz * ry) * rx;
    }
};

enum class RayType { PrimaryRay, ShadowRay };
struct Ray {
    RayType rayType = RayType::PrimaryRay;
    Vec3f orig{ 0, 0, 0 };
    Vec3f dir{ 0, 0, -1 };
};
