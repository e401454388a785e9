/*! \file eigen_shim.hpp
 *  \brief The handful of Eigen types that appear in the PUBLIC state of the reference's ICPStep / ICP classes
 *         (`Eigen::Matrix3f Rk, R; Eigen::Quaternionf qk, q; Eigen::Vector3f tk, t`,
 *         /root/reference/include/ICP/algorithms.hpp:1682-1697, :2302-2320), with the members the reference's callers
 *         use on them (/root/reference/src/ocl_icp_reg.cpp:190-205, /root/reference/src/ocl_icp_sbs.cpp:206-217):
 *
 *             reg.q.vec ().norm ()          reg.q.w ()          reg.q.vec ().normalized ()
 *             Eigen::Vector3f axis (cond ? Eigen::Vector3f::Zero () : ...)
 *             std::cout << axis.transpose () << reg.t.transpose ()          icpStep.tk.norm ()
 *
 *  so that those translation units compile unchanged against include/ICP/algorithms.hpp without Eigen installed.
 *  All rotation arithmetic runs on the device (icp_solve.cuh); nothing here is on the hot path.  If the real Eigen
 *  has been included first (EIGEN_CORE_H), this header defines nothing and the genuine types are used.
 */
#ifndef ICP_EIGEN_SHIM_HPP
#define ICP_EIGEN_SHIM_HPP

#ifndef EIGEN_CORE_H

#include <cmath>
#include <ostream>

namespace Eigen
{
    struct Vector3f;

    /*! \brief What `v.transpose ()` returns: a row view that only knows how to print itself (Eigen prints a row vector
     *         as its coefficients separated by one blank, the format the reference's log lines rely on). */
    struct RowVector3fView
    {
        float v[3];
        friend std::ostream& operator<< (std::ostream &os, const RowVector3fView &r) { return os << r.v[0] << " " << r.v[1] << " " << r.v[2]; }
    };

    struct Vector3f
    {
        float v[3] = { 0.f, 0.f, 0.f };
        Vector3f () {}
        Vector3f (float x_, float y_, float z_) { v[0] = x_; v[1] = y_; v[2] = z_; }
        static Vector3f Zero () { return Vector3f (); }
        float& operator[] (int i) { return v[i]; }
        float operator[] (int i) const { return v[i]; }
        float& operator() (int i) { return v[i]; }
        float operator() (int i) const { return v[i]; }
        float x () const { return v[0]; } float y () const { return v[1]; } float z () const { return v[2]; }
        const float* data () const { return v; }
        float squaredNorm () const { return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; }
        float norm () const { return std::sqrt (squaredNorm ()); }
        Vector3f normalized () const
        {
            const float n = norm ();
            return n > 0.f ? Vector3f (v[0] / n, v[1] / n, v[2] / n) : *this;       // Eigen leaves a zero vector unchanged
        }
        RowVector3fView transpose () const { RowVector3fView r; r.v[0] = v[0]; r.v[1] = v[1]; r.v[2] = v[2]; return r; }
        Vector3f operator+ (const Vector3f &o) const { return Vector3f (v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
        Vector3f operator- (const Vector3f &o) const { return Vector3f (v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
        Vector3f operator* (float s) const { return Vector3f (v[0] * s, v[1] * s, v[2] * s); }
        /*! column vector: one coefficient per line, like Eigen */
        friend std::ostream& operator<< (std::ostream &os, const Vector3f &a) { return os << a.v[0] << "\n" << a.v[1] << "\n" << a.v[2]; }
    };

    struct Matrix3f
    {
        float m[9] = { 1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f };   // row major (the C ABI's icp_state layout)
        static Matrix3f Identity () { return Matrix3f (); }
        float& operator() (int r, int c) { return m[r * 3 + c]; }
        float operator() (int r, int c) const { return m[r * 3 + c]; }
        Vector3f operator* (const Vector3f &x) const
        {
            return Vector3f (m[0] * x[0] + m[1] * x[1] + m[2] * x[2], m[3] * x[0] + m[4] * x[1] + m[5] * x[2], m[6] * x[0] + m[7] * x[1] + m[8] * x[2]);
        }
        Matrix3f transpose () const { Matrix3f t; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) t (r, c) = (*this) (c, r); return t; }
        float determinant () const
        {
            return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
        }
        friend std::ostream& operator<< (std::ostream &os, const Matrix3f &a)
        {
            for (int r = 0; r < 3; ++r) os << a (r, 0) << " " << a (r, 1) << " " << a (r, 2) << (r < 2 ? "\n" : "");
            return os;
        }
    };

    struct Quaternionf
    {
        float c[4] = { 0.f, 0.f, 0.f, 1.f };            // x y z w, the order of Eigen's coeffs ()
        Quaternionf () {}
        Quaternionf (float w_, float x_, float y_, float z_) { c[0] = x_; c[1] = y_; c[2] = z_; c[3] = w_; }    // Eigen's (w, x, y, z) constructor
        static Quaternionf Identity () { return Quaternionf (); }
        float x () const { return c[0]; } float y () const { return c[1]; } float z () const { return c[2]; } float w () const { return c[3]; }
        const float* coeffs () const { return c; }
        Vector3f vec () const { return Vector3f (c[0], c[1], c[2]); }
        float norm () const { return std::sqrt (c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + c[3] * c[3]); }
        /*! R (q), the formula of ICPTransform / Eigen's toRotationMatrix () */
        Matrix3f toRotationMatrix () const
        {
            const float xx = c[0] * c[0], yy = c[1] * c[1], zz = c[2] * c[2], xy = c[0] * c[1], xz = c[0] * c[2], yz = c[1] * c[2];
            const float wx = c[3] * c[0], wy = c[3] * c[1], wz = c[3] * c[2];
            Matrix3f r;
            r (0, 0) = 1.f - 2.f * (yy + zz); r (0, 1) = 2.f * (xy - wz);       r (0, 2) = 2.f * (xz + wy);
            r (1, 0) = 2.f * (xy + wz);       r (1, 1) = 1.f - 2.f * (xx + zz); r (1, 2) = 2.f * (yz - wx);
            r (2, 0) = 2.f * (xz - wy);       r (2, 1) = 2.f * (yz + wx);       r (2, 2) = 1.f - 2.f * (xx + yy);
            return r;
        }
    };
}

#endif  // EIGEN_CORE_H
#endif  // ICP_EIGEN_SHIM_HPP
