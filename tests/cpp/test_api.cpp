// test_api.cpp -- exercises the C++ drop-in class API (include/ICP/algorithms.hpp) the way the reference's gtests
// (/root/reference/tests/testsICP.cpp) use the OpenCL classes: construct . init . fill the staging buffer . write .
// run . read . compare with the CPU oracle.  Built and run by tests/test_cpp_api_gpu.py (GPU box only).
#include <ICP/algorithms.hpp>

#include <cmath>
#include <cstdio>
#include <iostream>
#include <random>
#include <vector>

using namespace cl_algo::ICP;

// CPU oracle (test infrastructure; oracle/libicp_oracle.so)
extern "C" {
void orc_get_lms(const float *, float *);
void orc_get_reps(const float *, uint32_t, uint32_t, uint32_t, float *);
void orc_transform_q(const float *, uint32_t, const float *, float *);
void orc_weights(const float *, uint32_t, float *, double *);
void orc_mean_weighted(const float *, const float *, const float *, double, uint32_t, float *);
void orc_devs(const float *, const float *, const float *, uint32_t, float *, float *);
void orc_sij(const float *, const float *, const float *, uint32_t, float, float *);
int orc_power_method(const float *, const float *, float *);
int orc_icp_register(const float *, const float *, uint32_t, uint32_t, uint32_t, uint32_t, float, float, int, int, int, uint32_t,
                     double, double, const float *, float *, float *, void *);
}

static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { ++failures; std::printf ("FAIL %s:%d: ", __FILE__, __LINE__); std::printf (__VA_ARGS__); std::printf ("\n"); } } while (0)

static bool same_bits (const float *a, const float *b, size_t n)
{
    return std::memcmp (a, b, n * sizeof (float)) == 0;
}

int main ()
{
    const unsigned int m = 16384, nr = 256, d = 8;
    std::mt19937 gen (1234);
    std::uniform_real_distribution<float> U01 (0.f, 1.f), U255 (0.f, 255.f);

    clutils::CLEnv clEnv;
    clEnv.addContext (0);
    clEnv.addQueue (0, 0, CL_QUEUE_PROFILING_ENABLE);
    clutils::CLEnvInfo<1> info (0, 0, 0, { 0 }, 0);

    {   // ICP.getLMs (testsICP.cpp:66-141)
        ICPLMs lm (clEnv, info);
        lm.init ();
        for (size_t i = 0; i < 640u * 480u * d; ++i) lm.hPtrIn[i] = U01 (gen);
        lm.write ();
        lm.run ();
        cl_float *res = (cl_float *) lm.read ();
        std::vector<float> ref (m * d);
        orc_get_lms (lm.hPtrIn, ref.data ());
        CHECK (same_bits (res, ref.data (), m * d), "getLMs");
    }
    {   // ICP.getReps (testsICP.cpp:147-222)
        ICPReps reps (clEnv, info);
        reps.init (nr);
        for (size_t i = 0; i < (size_t) m * d; ++i) reps.hPtrIn[i] = U01 (gen);
        reps.write ();
        reps.run ();
        cl_float *res = (cl_float *) reps.read ();
        std::vector<float> ref (nr * d);
        orc_get_reps (reps.hPtrIn, 128, 128, nr, ref.data ());
        CHECK (same_bits (res, ref.data (), nr * d), "getReps");
    }
    {   // ICP.icpTransform_Quaternion (testsICP.cpp:796-884)
        ICPTransform<ICPTransformConfig::QUATERNION> tr (clEnv, info);
        tr.init (m);
        for (size_t i = 0; i < (size_t) m * d; ++i) tr.hPtrInM[i] = U255 (gen);
        const float q[4] = { 0.5144f, 0.5743f, 0.5632f, 0.2973f };
        for (int i = 0; i < 4; ++i) tr.hPtrInT[i] = q[i];
        for (int i = 4; i < 7; ++i) tr.hPtrInT[i] = U255 (gen);
        tr.hPtrInT[7] = U01 (gen);
        tr.write (ICPTransform<ICPTransformConfig::QUATERNION>::Memory::D_IN_M);
        tr.write (ICPTransform<ICPTransformConfig::QUATERNION>::Memory::D_IN_T);
        tr.run ();
        cl_float *res = (cl_float *) tr.read ();
        std::vector<float> ref (m * d);
        orc_transform_q (tr.hPtrInM, m, tr.hPtrInT, ref.data ());
        CHECK (same_bits (res, ref.data (), m * d), "transformQ");
    }
    std::vector<float> W (m);
    double sumW = 0.0;
    {   // ICP.icpComputeReduceWeights (testsICP.cpp:228-321)
        ICPWeights w (clEnv, info);
        w.init (m);
        std::vector<float> dist (m);
        for (unsigned i = 0; i < m; ++i) { dist[i] = 1e-6f + 254e-6f * U01 (gen); w.hPtrIn[i].dist = dist[i]; w.hPtrIn[i].id = i; }
        w.write ();
        w.run ();
        cl_float *rw = (cl_float *) w.read (ICPWeights::Memory::H_OUT_W, CL_FALSE);
        cl_double *rs = (cl_double *) w.read (ICPWeights::Memory::H_OUT_SUM_W);
        orc_weights (dist.data (), m, W.data (), &sumW);
        CHECK (same_bits (rw, W.data (), m), "weights");
        CHECK (*rs == sumW, "sum of weights %.17g vs %.17g", *rs, sumW);
    }
    std::vector<float> F (m * d), M (m * d), mean (8);
    {   // ICP.icpMean_Weighted (testsICP.cpp:412-498)
        ICPMean<ICPMeanConfig::WEIGHTED> mn (clEnv, info);
        mn.init (m);
        for (size_t i = 0; i < (size_t) m * d; ++i) { F[i] = mn.hPtrInF[i] = U01 (gen); M[i] = mn.hPtrInM[i] = U01 (gen); }
        std::memcpy (mn.hPtrInW, W.data (), m * sizeof (float));
        mn.hPtrInSW[0] = sumW;
        typedef ICPMean<ICPMeanConfig::WEIGHTED>::Memory Mem;
        mn.write (Mem::D_IN_F); mn.write (Mem::D_IN_M); mn.write (Mem::D_IN_W); mn.write (Mem::D_IN_SUM_W);
        mn.run ();
        cl_float *res = (cl_float *) mn.read ();
        orc_mean_weighted (F.data (), M.data (), W.data (), sumW, m, mean.data ());
        CHECK (same_bits (res, mean.data (), 8), "mean_weighted");
    }
    std::vector<float> DF (m * 4), DM (m * 4);
    {   // ICP.icpSubtractMean (testsICP.cpp:504-595) chained into ICP.icpSijProducts_Weighted (:696-789) by buffer sharing
        ICPDevs dv (clEnv, info);
        dv.init (m);
        std::memcpy (dv.hPtrInF, F.data (), F.size () * sizeof (float));
        std::memcpy (dv.hPtrInM, M.data (), M.size () * sizeof (float));
        std::memcpy (dv.hPtrInMean, mean.data (), 8 * sizeof (float));
        dv.write (ICPDevs::Memory::D_IN_F); dv.write (ICPDevs::Memory::D_IN_M); dv.write (ICPDevs::Memory::D_IN_MEAN);
        ICPS<ICPSConfig::WEIGHTED> sm (clEnv, info);
        typedef ICPS<ICPSConfig::WEIGHTED>::Memory SM;
        sm.get (SM::D_IN_DEV_F) = dv.get (ICPDevs::Memory::D_OUT_DEV_F);     // assigned before init (): shared, not re-created
        sm.get (SM::D_IN_DEV_M) = dv.get (ICPDevs::Memory::D_OUT_DEV_M);
        sm.init (m, 1e-6f);
        CHECK (sm.get (SM::D_IN_DEV_F) () == dv.get (ICPDevs::Memory::D_OUT_DEV_F) (), "buffer sharing through get ()");
        std::memcpy (sm.hPtrInW, W.data (), m * sizeof (float));
        sm.write (SM::D_IN_W);
        dv.run ();
        sm.run ();
        cl_float *rdf = (cl_float *) dv.read (ICPDevs::Memory::H_OUT_DEV_F, CL_FALSE);
        cl_float *rdm = (cl_float *) dv.read (ICPDevs::Memory::H_OUT_DEV_M);
        cl_float *rs = (cl_float *) sm.read ();
        orc_devs (F.data (), M.data (), mean.data (), m, DF.data (), DM.data ());
        float S[11];
        orc_sij (DM.data (), DF.data (), W.data (), m, 1e-6f, S);
        CHECK (same_bits (rdf, DF.data (), m * 4) && same_bits (rdm, DM.data (), m * 4), "devs");
        CHECK (same_bits (rs, S, 11), "Sij weighted");
    }
    {   // ICP.icpPowerMethod with the reference's known-answer vector (testsICP.cpp:988-1087)
        ICPPowerMethod pm (clEnv, info);
        pm.init ();
        const cl_float S[11] = { 0.00168053f, 0.000131408f, -0.000775179f, 0.000156595f, 0.00102674f, -0.000563479f,
                                 -0.000722137f, -0.000559463f, 0.00246661f, 0.00521271f, 0.00515292f };
        const cl_float means[8] = { -33.9694f, -17.6421f, 1494.22f, 0.f, -44.8322f, -19.3835f, 1485.93f, 0.f };
        pm.write (ICPPowerMethod::Memory::D_IN_S, (void *) S);
        pm.write (ICPPowerMethod::Memory::D_IN_MEAN, (void *) means);
        pm.run ();
        cl_float *res = (cl_float *) pm.read ();
        float refTk[8];
        orc_power_method (S, means, refTk);
        CHECK (same_bits (res, refTk, 8), "power method vs CPU");
        const cl_float svdTk[8] = { 0.00111412f, 0.00730956f, -0.00647493f, 0.999952f, -10.4598f, 4.74009f, -0.762817f, 1.00578f };
        const float eps = 42000 * 1.1920929e-07f;
        for (int k = 0; k < 8; ++k) CHECK (std::fabs (svdTk[k] - res[k]) < eps, "power method vs SVD golden, element %d", k);
    }
    {   // registration as ICPReg does it (src/ocl_icp_reg.cpp:81-121,165-175): ICP<POWER_METHOD,WEIGHTED>
        // F = smooth synthetic surface, M = F moved by a small rigid motion
        std::vector<float> Fx (m * d), Mx (m * d);
        for (unsigned i = 0; i < m; ++i)
        {
            const float u = (float) (i % 128) - 64.f, v = (float) (i / 128) - 64.f;
            const float z = 1500.f + 0.8f * u + 40.f * std::sin (0.07f * v) + 25.f * std::cos (0.05f * u);
            float *f = &Fx[i * d];
            f[0] = u * z / 595.f * 4.f; f[1] = v * z / 595.f * 3.f; f[2] = z; f[3] = 1.f;
            f[4] = 0.5f + 0.4f * std::sin (0.11f * u + 0.05f * v); f[5] = 0.5f + 0.4f * std::cos (0.09f * v); f[6] = 0.5f + 0.3f * std::sin (0.05f * (u + v)); f[7] = 1.f;
            float *g = &Mx[i * d];
            const float th = 0.02f;
            g[0] = std::cos (th) * f[0] - std::sin (th) * f[1] + 6.f; g[1] = std::sin (th) * f[0] + std::cos (th) * f[1] - 4.f; g[2] = f[2] + 3.f; g[3] = 1.f;
            g[4] = f[4]; g[5] = f[5]; g[6] = f[6]; g[7] = 1.f;
        }
        typedef ICP<ICPStepConfigT::POWER_METHOD, ICPStepConfigW::WEIGHTED> Reg;
        Reg reg (clEnv, info, info);
        reg.init (m, nr, 2e2f, 1e-6f, 40, 0.001, 0.01);
        reg.write (Reg::Memory::D_IN_F, Fx.data ());
        reg.write (Reg::Memory::D_IN_M, Mx.data ());
        reg.buildRBC ();
        reg.run ();
        float T[8], T16[16];
        const int k = orc_icp_register (Fx.data (), Mx.data (), m, 128, 128, nr, 2e2f, 1e-6f, 1, 1, 0, 40, 0.001, 0.01, nullptr, T, T16, nullptr);
        CHECK ((int) reg.k == k, "ICP::k %u vs oracle %d", reg.k, k);
        cl_float *res = (cl_float *) reg.read ();
        CHECK (same_bits (res, T, 8), "ICP pose {q,t,s}");
        CHECK (reg.q.x () == T[0] && reg.q.w () == T[3] && reg.t[0] == T[4] && reg.s == T[7], "public members q, t, s");
        for (int r = 0; r < 3; ++r) for (int c2 = 0; c2 < 3; ++c2)
            CHECK (std::fabs (reg.s * reg.R (r, c2) - T16[r * 4 + c2]) <= 1e-5f, "4x4 pose element (%d,%d)", r, c2);
        std::printf ("registration: %u iterations, t = (%.3f, %.3f, %.3f), s = %.5f\n", reg.k, reg.t[0], reg.t[1], reg.t[2], reg.s);

        {   // The print block of ICPReg<RC,WC>::registerPC, VERBATIM from the reference (src/ocl_icp_reg.cpp:190-205; only the
            // timer line is dropped): it must compile unchanged against the drop-in header -- q.vec ().norm (), .normalized (),
            // Eigen::Vector3f::Zero (), .transpose (), operator<<.
            double sinth_2 = reg.q.vec ().norm ();
            double angle = 180.f / M_PI * 2 * std::atan2 (sinth_2, reg.q.w ());
            Eigen::Vector3f axis ((sinth_2 == 0.0) ? Eigen::Vector3f::Zero () : reg.q.vec ().normalized ());

            std::cout << std::endl << "================" << std::endl << std::endl;
            std::cout << "    Iterations            :    " << reg.k << std::endl;
            std::cout << "    Rotation angle        :    " << angle << " degrees" << std::endl;
            std::cout << "    Rotation axis         :    " << axis.transpose () << std::endl;
            std::cout << "    Translation vector    :    " << reg.t.transpose () << std::endl;
            std::cout << "    Scale                 :    " << reg.s << std::endl;
            CHECK (std::fabs (axis.norm () - 1.f) < 1e-5f || sinth_2 == 0.0, "rotation axis is not a unit vector");
            CHECK (angle > 0.0 && angle < 5.0, "rotation angle %.4f degrees is not a small positive angle", angle);
        }

        // step by step (ICPSBS::step, src/ocl_icp_sbs.cpp:167-218): same result as 5 fixed iterations of the oracle
        typedef ICPStep<ICPStepConfigT::EIGEN, ICPStepConfigW::REGULAR> Step;
        Step st (clEnv, info, info);
        st.init (m, nr, 2e2f, 1e-6f);
        st.write (Step::Memory::D_IN_F, Fx.data ());
        st.write (Step::Memory::D_IN_M, Mx.data ());
        st.buildRBC ();
        st.run (nullptr, nullptr, true);
        for (int i = 1; i < 5; ++i) st.run ();
        orc_icp_register (Fx.data (), Mx.data (), m, 128, 128, nr, 2e2f, 1e-6f, 0, 0, 5, 40, 0.001, 0.01, nullptr, T, T16, nullptr);
        res = (cl_float *) st.read ();
        CHECK (same_bits (res, T, 8), "ICPStep<EIGEN,REGULAR> pose after 5 steps");
        {   // the extra lines of ICPSBS<RC,WC>::step's print block (src/ocl_icp_sbs.cpp:215-217), verbatim
            Step &icpStep = st;
            std::cout << "    Change in translation :    " << icpStep.tk.norm () << " mm" << std::endl;
            std::cout << "    Change in rotation    :    " << 180.f / M_PI * 2 * std::atan2 (
                icpStep.qk.vec ().norm (), icpStep.qk.w ()) << " degrees" << std::endl << std::endl;
            // R is the accumulated rotation, q the same rotation as a quaternion (algorithms.cpp:4688-4689)
            const Eigen::Matrix3f Rq = icpStep.q.toRotationMatrix ();
            for (int r = 0; r < 3; ++r) for (int c2 = 0; c2 < 3; ++c2)
                CHECK (std::fabs (Rq (r, c2) - icpStep.R (r, c2)) <= 2e-6f, "R vs R(q) element (%d,%d)", r, c2);
        }
    }
    {   // ICPBatch<CR,CW>: independent pairs per call over every visible GPU (icp_multi_*); pose p = what ICP<CR,CW>'s fixed
        // 40-step profiling driver (algorithms.hpp:2482-2494) computes for pair p
        const unsigned n_pairs = 3, iters = 8;
        typedef ICPBatch<ICPStepConfigT::POWER_METHOD, ICPStepConfigW::WEIGHTED> Batch;
        Batch batch (clEnv, info, info);
        batch.init (n_pairs, m, nr, 2e2f, 1e-6f, iters, 0);
        std::vector<float> Fx (m * d), Mx (m * d);
        for (unsigned p = 0; p < n_pairs; ++p)
        {
            const float th = 0.01f * (float) (p + 1), tx = 3.f + 2.f * (float) p;
            for (unsigned i = 0; i < m; ++i)
            {
                const float u = (float) (i % 128) - 64.f, v = (float) (i / 128) - 64.f;
                const float z = 1400.f + 0.6f * u + 35.f * std::sin (0.06f * v + (float) p) + 20.f * std::cos (0.05f * u);
                float *f = &Fx[i * d];
                f[0] = u * z / 595.f * 4.f; f[1] = v * z / 595.f * 3.f; f[2] = z; f[3] = 1.f;
                f[4] = 0.5f + 0.4f * std::sin (0.11f * u + 0.05f * v); f[5] = 0.5f + 0.4f * std::cos (0.09f * v); f[6] = 0.5f + 0.3f * std::sin (0.05f * (u + v)); f[7] = 1.f;
                float *g = &Mx[i * d];
                g[0] = std::cos (th) * f[0] - std::sin (th) * f[1] + tx; g[1] = std::sin (th) * f[0] + std::cos (th) * f[1] - 2.f; g[2] = f[2] + 1.5f; g[3] = 1.f;
                g[4] = f[4]; g[5] = f[5]; g[6] = f[6]; g[7] = 1.f;
            }
            batch.write (Batch::Memory::H_IN_F, p, Fx.data ());
            batch.write (Batch::Memory::H_IN_M, p, Mx.data ());
        }
        batch.run ();
        CHECK (batch.devices () >= 1, "ICPBatch uses no device");
        for (unsigned p = 0; p < n_pairs; ++p)
        {
            float T[8], T16[16];
            orc_icp_register (batch.hPtrInF + (size_t) p * m * d, batch.hPtrInM + (size_t) p * m * d, m, 128, 128, nr, 2e2f, 1e-6f, 1, 1, (int) iters, 40, 0.001, 0.01,
                              nullptr, T, T16, nullptr);
            CHECK (same_bits (batch.read (p), T, 8), "ICPBatch pose of pair %u", p);
        }
        std::printf ("batch: %u pairs over %d device(s)\n", n_pairs, batch.devices ());
    }
    if (failures == 0) std::printf ("ALL C++ API CHECKS PASSED\n");
    return failures == 0 ? 0 : 1;
}
