/*! \file algorithms.hpp
 *  \brief Drop-in counterpart of the reference's algorithm-class API (cl_algo::ICP::*,
 *         /root/reference/include/ICP/algorithms.hpp) on top of libicp_b200.so (include/icp_b200.h).
 *
 *  Same class names, `Memory` enumerators, method shapes (`get` . `init` . `write` . `read` . `run` . timed `run`),
 *  public staging pointers (`hPtr*`) and ownership rule: a buffer assigned through `get (Memory)` BEFORE `init ()`
 *  is kept, `init ()` only creates what is still null (algorithms.cpp:216-221).  Device code = hand written sm_100a
 *  CUDA kernels; there is no OpenCL and no CPU fallback.  Header only: every method forwards to the C ABI.
 *
 *  Differences a maintainer should know (all documented in INTEGRATION.md):
 *   - `cl::` / `clutils::` types are the stand-ins of ICP/cl_shim.hpp (CUDA device memory / stream underneath);
 *   - `ICPStep` keeps the registration state on the device; its public members Rk, qk, tk, sk, R, q, t, s are
 *     refreshed by `run ()` (one blocking 136-byte read, like the reference's blocking read of Tk) and the whole
 *     `ICP::run ()` loop, including ICP::check (), executes on the device without host round trips;
 *   - Eigen is not required: ICP/eigen_shim.hpp provides `Eigen::Matrix3f / Quaternionf / Vector3f` look-alikes with the
 *     members the reference's callers use (`vec ()`, `norm ()`, `normalized ()`, `transpose ()`, `Zero ()`, `<<`), so
 *     `src/ocl_icp_reg.cpp:190-205` compiles unchanged; include the real Eigen first to use the genuine types.
 */
#ifndef ICP_ALGORITHMS_HPP
#define ICP_ALGORITHMS_HPP

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "common.hpp"
#include "cl_shim.hpp"
#include "eigen_shim.hpp"

namespace cl_algo
{
namespace ICP
{
    enum class ReduceConfig : uint8_t { MIN, MAX, SUM };                       // algorithms.hpp:52-57
    enum class ScanConfig : uint8_t { INCLUSIVE, EXCLUSIVE };                  // :189-193
    enum class ICPMeanConfig : uint8_t { REGULAR, WEIGHTED };                  // :577-581
    enum class ICPSConfig : uint8_t { REGULAR, WEIGHTED };                     // :944-948
    enum class ICPTransformConfig : uint8_t { QUATERNION, MATRIX };            // :1190-1194
    enum class ICPStepConfigT : uint8_t { EIGEN, POWER_METHOD, JACOBI };       // :1544-1556
    enum class ICPStepConfigW : uint8_t { REGULAR, WEIGHTED };                 // :1560-1564

    // ---- the Eigen types of ICPStep's public state (algorithms.hpp:2302-2320): the real ones when Eigen was included
    //      before this header, otherwise the minimal look-alikes of ICP/eigen_shim.hpp (same accessors the reference's
    //      callers use: src/ocl_icp_reg.cpp:190-205, src/ocl_icp_sbs.cpp:206-217) ----
    using Eigen::Vector3f;
    using Eigen::Matrix3f;
    using Eigen::Quaternionf;

    namespace detail
    {
        /*! \brief Slot table shared by all stage classes: one cl::Buffer per `Memory` enumerator. */
        template <unsigned N>
        struct Slots
        {
            cl::Buffer buf[N];
            size_t bytes[N] = {};
            cl::Memory& at (unsigned i) { return buf[i]; }
            /*! Create the buffer unless the caller assigned one through get () before init (). */
            void ensure (const cl::Context &ctx, unsigned i, size_t nbytes, bool host)
            {
                bytes[i] = nbytes;
                if (buf[i] () == nullptr)
                    buf[i] = cl::Buffer (ctx, host ? CL_MEM_ALLOC_HOST_PTR : CL_MEM_READ_WRITE, nbytes ? nbytes : 16);
            }
            template <typename T> T* host (unsigned i) { return reinterpret_cast<T *> (buf[i] ()); }
            void* dev (unsigned i) { return buf[i] (); }
        };
        inline bool stage_in (Staging s) { return s == Staging::I || s == Staging::IO; }
        inline bool stage_out (Staging s) { return s == Staging::O || s == Staging::IO; }
    }

    /*! \brief ICPLMs stage (algorithms.hpp:312-374). */
    class ICPLMs
    {
    public:
        enum class Memory : uint8_t { H_IN, H_OUT, D_IN, D_OUT };
        ICPLMs (clutils::CLEnv &_env, clutils::CLEnvInfo<1> _info) : env (_env), info (_info), ctx (_env.getContext (_info.ctxIdx)), queue (ctx) {}
        cl::Memory& get (Memory mem) { return S.at (idx (mem)); }
        void init (Staging _staging = Staging::IO)
        {
            staging = _staging;
            if (detail::stage_in (staging)) { S.ensure (ctx, idx (Memory::H_IN), 640u * 480u * 8u * sizeof (cl_float), true); hPtrIn = S.host<cl_float> (idx (Memory::H_IN)); } else hPtrIn = nullptr;
            S.ensure (ctx, idx (Memory::D_IN), 640u * 480u * 8u * sizeof (cl_float), false);
            if (detail::stage_out (staging)) { S.ensure (ctx, idx (Memory::H_OUT), 16384u * 8u * sizeof (cl_float), true); hPtrOut = S.host<cl_float> (idx (Memory::H_OUT)); } else hPtrOut = nullptr;
            S.ensure (ctx, idx (Memory::D_OUT), 16384u * 8u * sizeof (cl_float), false);
        }
        void write (Memory mem = Memory::D_IN, void *ptr = nullptr, bool block = CL_FALSE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_in (staging)) return;
            switch (mem)
            {
                case Memory::D_IN:
                    if (ptr != nullptr) std::memcpy (hPtrIn, ptr, S.bytes[idx (Memory::D_IN)]);
                    queue.enqueueWriteBuffer (S.at (idx (Memory::D_IN)), block, 0, S.bytes[idx (Memory::D_IN)], hPtrIn);
                    break;
                default: break;
            }
        }
        void* read (Memory mem = Memory::H_OUT, bool block = CL_TRUE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_out (staging)) return nullptr;
            switch (mem)
            {
                case Memory::H_OUT:
                    queue.enqueueReadBuffer (S.at (idx (Memory::D_OUT)), block, 0, S.bytes[idx (Memory::D_OUT)], hPtrOut);
                    return hPtrOut;
                default: return nullptr;
            }
        }
        void run (const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            cl::check (icp_get_lms (ctx (), (const float *) S.dev (idx (Memory::D_IN)), (float *) S.dev (idx (Memory::D_OUT))));
        }
        /*! \brief Timed execution (ms), the counterpart of run (clutils::GPUTimer&). */
        template <typename period>
        double run (clutils::GPUTimer<period> &timer, const std::vector<cl::Event> *events = nullptr)
        {
            timer.start (); run (events); return timer.stop ();
        }
        cl_float *hPtrIn = nullptr;  /*!< mapping of the H_IN staging buffer */
        cl_float *hPtrOut = nullptr;  /*!< mapping of the H_OUT staging buffer */
    private:
        static unsigned idx (Memory m) { return static_cast<unsigned> (m); }
        clutils::CLEnv &env;
        clutils::CLEnvInfo<1> info;
        cl::Context ctx;
        cl::CommandQueue queue;
        Staging staging = Staging::IO;
        detail::Slots<4> S;
    };

    /*! \brief ICPReps stage (algorithms.hpp:397-459). */
    class ICPReps
    {
    public:
        enum class Memory : uint8_t { H_IN, H_OUT, D_IN, D_OUT };
        ICPReps (clutils::CLEnv &_env, clutils::CLEnvInfo<1> _info) : env (_env), info (_info), ctx (_env.getContext (_info.ctxIdx)), queue (ctx) {}
        cl::Memory& get (Memory mem) { return S.at (idx (mem)); }
        void init (unsigned int _nr, Staging _staging = Staging::IO)
        {
            staging = _staging;
            nr = _nr;
            if (detail::stage_in (staging)) { S.ensure (ctx, idx (Memory::H_IN), 128u * 128u * 8u * sizeof (cl_float), true); hPtrIn = S.host<cl_float> (idx (Memory::H_IN)); } else hPtrIn = nullptr;
            S.ensure (ctx, idx (Memory::D_IN), 128u * 128u * 8u * sizeof (cl_float), false);
            if (detail::stage_out (staging)) { S.ensure (ctx, idx (Memory::H_OUT), nr * 8u * sizeof (cl_float), true); hPtrOut = S.host<cl_float> (idx (Memory::H_OUT)); } else hPtrOut = nullptr;
            S.ensure (ctx, idx (Memory::D_OUT), nr * 8u * sizeof (cl_float), false);
        }
        void write (Memory mem = Memory::D_IN, void *ptr = nullptr, bool block = CL_FALSE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_in (staging)) return;
            switch (mem)
            {
                case Memory::D_IN:
                    if (ptr != nullptr) std::memcpy (hPtrIn, ptr, S.bytes[idx (Memory::D_IN)]);
                    queue.enqueueWriteBuffer (S.at (idx (Memory::D_IN)), block, 0, S.bytes[idx (Memory::D_IN)], hPtrIn);
                    break;
                default: break;
            }
        }
        void* read (Memory mem = Memory::H_OUT, bool block = CL_TRUE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_out (staging)) return nullptr;
            switch (mem)
            {
                case Memory::H_OUT:
                    queue.enqueueReadBuffer (S.at (idx (Memory::D_OUT)), block, 0, S.bytes[idx (Memory::D_OUT)], hPtrOut);
                    return hPtrOut;
                default: return nullptr;
            }
        }
        void run (const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            cl::check (icp_get_reps (ctx (), (const float *) S.dev (idx (Memory::D_IN)), 128, 128, nr, (float *) S.dev (idx (Memory::D_OUT))));
        }
        /*! \brief Timed execution (ms), the counterpart of run (clutils::GPUTimer&). */
        template <typename period>
        double run (clutils::GPUTimer<period> &timer, const std::vector<cl::Event> *events = nullptr)
        {
            timer.start (); run (events); return timer.stop ();
        }
        cl_float *hPtrIn = nullptr;  /*!< mapping of the H_IN staging buffer */
        cl_float *hPtrOut = nullptr;  /*!< mapping of the H_OUT staging buffer */
    protected:
        unsigned int nr = 0;
    private:
        static unsigned idx (Memory m) { return static_cast<unsigned> (m); }
        clutils::CLEnv &env;
        clutils::CLEnvInfo<1> info;
        cl::Context ctx;
        cl::CommandQueue queue;
        Staging staging = Staging::IO;
        detail::Slots<4> S;
    };

    /*! \brief ICPWeights stage (algorithms.hpp:485-572). */
    class ICPWeights
    {
    public:
        enum class Memory : uint8_t { H_IN, H_OUT_W, H_OUT_SUM_W, D_IN, D_OUT_W, D_GW, D_OUT_SUM_W };
        ICPWeights (clutils::CLEnv &_env, clutils::CLEnvInfo<1> _info) : env (_env), info (_info), ctx (_env.getContext (_info.ctxIdx)), queue (ctx) {}
        cl::Memory& get (Memory mem) { return S.at (idx (mem)); }
        void init (unsigned int _n, Staging _staging = Staging::IO)
        {
            staging = _staging;
            n = _n;
            if (detail::stage_in (staging)) { S.ensure (ctx, idx (Memory::H_IN), n * sizeof (rbc_dist_id), true); hPtrIn = S.host<rbc_dist_id> (idx (Memory::H_IN)); } else hPtrIn = nullptr;
            S.ensure (ctx, idx (Memory::D_IN), n * sizeof (rbc_dist_id), false);
            if (detail::stage_out (staging)) { S.ensure (ctx, idx (Memory::H_OUT_W), n * sizeof (cl_float), true); hPtrOutW = S.host<cl_float> (idx (Memory::H_OUT_W)); } else hPtrOutW = nullptr;
            S.ensure (ctx, idx (Memory::D_OUT_W), n * sizeof (cl_float), false);
            if (detail::stage_out (staging)) { S.ensure (ctx, idx (Memory::H_OUT_SUM_W), sizeof (cl_double), true); hPtrOutSW = S.host<cl_double> (idx (Memory::H_OUT_SUM_W)); } else hPtrOutSW = nullptr;
            S.ensure (ctx, idx (Memory::D_OUT_SUM_W), sizeof (cl_double), false);
        }
        void write (Memory mem = Memory::D_IN, void *ptr = nullptr, bool block = CL_FALSE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_in (staging)) return;
            switch (mem)
            {
                case Memory::D_IN:
                    if (ptr != nullptr) std::memcpy (hPtrIn, ptr, S.bytes[idx (Memory::D_IN)]);
                    queue.enqueueWriteBuffer (S.at (idx (Memory::D_IN)), block, 0, S.bytes[idx (Memory::D_IN)], hPtrIn);
                    break;
                default: break;
            }
        }
        void* read (Memory mem = Memory::H_OUT_SUM_W, bool block = CL_TRUE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_out (staging)) return nullptr;
            switch (mem)
            {
                case Memory::H_OUT_W:
                    queue.enqueueReadBuffer (S.at (idx (Memory::D_OUT_W)), block, 0, S.bytes[idx (Memory::D_OUT_W)], hPtrOutW);
                    return hPtrOutW;
                case Memory::H_OUT_SUM_W:
                    queue.enqueueReadBuffer (S.at (idx (Memory::D_OUT_SUM_W)), block, 0, S.bytes[idx (Memory::D_OUT_SUM_W)], hPtrOutSW);
                    return hPtrOutSW;
                default: return nullptr;
            }
        }
        void run (const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            cl::check (icp_weights (ctx (), (const icp_dist_id *) S.dev (idx (Memory::D_IN)), (float *) S.dev (idx (Memory::D_OUT_W)), (double *) S.dev (idx (Memory::D_OUT_SUM_W)), n));
        }
        /*! \brief Timed execution (ms), the counterpart of run (clutils::GPUTimer&). */
        template <typename period>
        double run (clutils::GPUTimer<period> &timer, const std::vector<cl::Event> *events = nullptr)
        {
            timer.start (); run (events); return timer.stop ();
        }
        rbc_dist_id *hPtrIn = nullptr;  /*!< mapping of the H_IN staging buffer */
        cl_float *hPtrOutW = nullptr;  /*!< mapping of the H_OUT_W staging buffer */
        cl_double *hPtrOutSW = nullptr;  /*!< mapping of the H_OUT_SUM_W staging buffer */
    protected:
        unsigned int n = 0;
    private:
        static unsigned idx (Memory m) { return static_cast<unsigned> (m); }
        clutils::CLEnv &env;
        clutils::CLEnvInfo<1> info;
        cl::Context ctx;
        cl::CommandQueue queue;
        Staging staging = Staging::IO;
        detail::Slots<7> S;
    };

    /*! \brief ICPDevs stage (algorithms.hpp:867-939). */
    class ICPDevs
    {
    public:
        enum class Memory : uint8_t { H_IN_F, H_IN_M, H_IN_MEAN, H_OUT_DEV_F, H_OUT_DEV_M, D_IN_F, D_IN_M, D_IN_MEAN, D_OUT_DEV_F, D_OUT_DEV_M };
        ICPDevs (clutils::CLEnv &_env, clutils::CLEnvInfo<1> _info) : env (_env), info (_info), ctx (_env.getContext (_info.ctxIdx)), queue (ctx) {}
        cl::Memory& get (Memory mem) { return S.at (idx (mem)); }
        void init (unsigned int _n, Staging _staging = Staging::IO)
        {
            staging = _staging;
            n = _n;
            if (detail::stage_in (staging)) { S.ensure (ctx, idx (Memory::H_IN_F), n * 8u * sizeof (cl_float), true); hPtrInF = S.host<cl_float> (idx (Memory::H_IN_F)); } else hPtrInF = nullptr;
            S.ensure (ctx, idx (Memory::D_IN_F), n * 8u * sizeof (cl_float), false);
            if (detail::stage_in (staging)) { S.ensure (ctx, idx (Memory::H_IN_M), n * 8u * sizeof (cl_float), true); hPtrInM = S.host<cl_float> (idx (Memory::H_IN_M)); } else hPtrInM = nullptr;
            S.ensure (ctx, idx (Memory::D_IN_M), n * 8u * sizeof (cl_float), false);
            if (detail::stage_in (staging)) { S.ensure (ctx, idx (Memory::H_IN_MEAN), 8u * sizeof (cl_float), true); hPtrInMean = S.host<cl_float> (idx (Memory::H_IN_MEAN)); } else hPtrInMean = nullptr;
            S.ensure (ctx, idx (Memory::D_IN_MEAN), 8u * sizeof (cl_float), false);
            if (detail::stage_out (staging)) { S.ensure (ctx, idx (Memory::H_OUT_DEV_F), n * 4u * sizeof (cl_float), true); hPtrOutDevF = S.host<cl_float> (idx (Memory::H_OUT_DEV_F)); } else hPtrOutDevF = nullptr;
            S.ensure (ctx, idx (Memory::D_OUT_DEV_F), n * 4u * sizeof (cl_float), false);
            if (detail::stage_out (staging)) { S.ensure (ctx, idx (Memory::H_OUT_DEV_M), n * 4u * sizeof (cl_float), true); hPtrOutDevM = S.host<cl_float> (idx (Memory::H_OUT_DEV_M)); } else hPtrOutDevM = nullptr;
            S.ensure (ctx, idx (Memory::D_OUT_DEV_M), n * 4u * sizeof (cl_float), false);
        }
        void write (Memory mem = Memory::D_IN_F, void *ptr = nullptr, bool block = CL_FALSE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_in (staging)) return;
            switch (mem)
            {
                case Memory::D_IN_F:
                    if (ptr != nullptr) std::memcpy (hPtrInF, ptr, S.bytes[idx (Memory::D_IN_F)]);
                    queue.enqueueWriteBuffer (S.at (idx (Memory::D_IN_F)), block, 0, S.bytes[idx (Memory::D_IN_F)], hPtrInF);
                    break;
                case Memory::D_IN_M:
                    if (ptr != nullptr) std::memcpy (hPtrInM, ptr, S.bytes[idx (Memory::D_IN_M)]);
                    queue.enqueueWriteBuffer (S.at (idx (Memory::D_IN_M)), block, 0, S.bytes[idx (Memory::D_IN_M)], hPtrInM);
                    break;
                case Memory::D_IN_MEAN:
                    if (ptr != nullptr) std::memcpy (hPtrInMean, ptr, S.bytes[idx (Memory::D_IN_MEAN)]);
                    queue.enqueueWriteBuffer (S.at (idx (Memory::D_IN_MEAN)), block, 0, S.bytes[idx (Memory::D_IN_MEAN)], hPtrInMean);
                    break;
                default: break;
            }
        }
        void* read (Memory mem = Memory::H_OUT_DEV_F, bool block = CL_TRUE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_out (staging)) return nullptr;
            switch (mem)
            {
                case Memory::H_OUT_DEV_F:
                    queue.enqueueReadBuffer (S.at (idx (Memory::D_OUT_DEV_F)), block, 0, S.bytes[idx (Memory::D_OUT_DEV_F)], hPtrOutDevF);
                    return hPtrOutDevF;
                case Memory::H_OUT_DEV_M:
                    queue.enqueueReadBuffer (S.at (idx (Memory::D_OUT_DEV_M)), block, 0, S.bytes[idx (Memory::D_OUT_DEV_M)], hPtrOutDevM);
                    return hPtrOutDevM;
                default: return nullptr;
            }
        }
        void run (const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            cl::check (icp_devs (ctx (), (const float *) S.dev (idx (Memory::D_IN_F)), (const float *) S.dev (idx (Memory::D_IN_M)), (const float *) S.dev (idx (Memory::D_IN_MEAN)), (float *) S.dev (idx (Memory::D_OUT_DEV_F)), (float *) S.dev (idx (Memory::D_OUT_DEV_M)), n));
        }
        /*! \brief Timed execution (ms), the counterpart of run (clutils::GPUTimer&). */
        template <typename period>
        double run (clutils::GPUTimer<period> &timer, const std::vector<cl::Event> *events = nullptr)
        {
            timer.start (); run (events); return timer.stop ();
        }
        cl_float *hPtrInF = nullptr;  /*!< mapping of the H_IN_F staging buffer */
        cl_float *hPtrInM = nullptr;  /*!< mapping of the H_IN_M staging buffer */
        cl_float *hPtrInMean = nullptr;  /*!< mapping of the H_IN_MEAN staging buffer */
        cl_float *hPtrOutDevF = nullptr;  /*!< mapping of the H_OUT_DEV_F staging buffer */
        cl_float *hPtrOutDevM = nullptr;  /*!< mapping of the H_OUT_DEV_M staging buffer */
    protected:
        unsigned int n = 0;
    private:
        static unsigned idx (Memory m) { return static_cast<unsigned> (m); }
        clutils::CLEnv &env;
        clutils::CLEnvInfo<1> info;
        cl::Context ctx;
        cl::CommandQueue queue;
        Staging staging = Staging::IO;
        detail::Slots<10> S;
    };

    /*! \brief ICPPowerMethod stage (algorithms.hpp:1451-1537). */
    class ICPPowerMethod
    {
    public:
        enum class Memory : uint8_t { H_IN_S, H_IN_MEAN, H_OUT_T_K, D_IN_S, D_IN_MEAN, D_OUT_T_K };
        ICPPowerMethod (clutils::CLEnv &_env, clutils::CLEnvInfo<1> _info) : env (_env), info (_info), ctx (_env.getContext (_info.ctxIdx)), queue (ctx) {}
        cl::Memory& get (Memory mem) { return S.at (idx (mem)); }
        void init (Staging _staging = Staging::IO)
        {
            staging = _staging;
            if (detail::stage_in (staging)) { S.ensure (ctx, idx (Memory::H_IN_S), 11u * sizeof (cl_float), true); hPtrInS = S.host<cl_float> (idx (Memory::H_IN_S)); } else hPtrInS = nullptr;
            S.ensure (ctx, idx (Memory::D_IN_S), 11u * sizeof (cl_float), false);
            if (detail::stage_in (staging)) { S.ensure (ctx, idx (Memory::H_IN_MEAN), 8u * sizeof (cl_float), true); hPtrInMean = S.host<cl_float> (idx (Memory::H_IN_MEAN)); } else hPtrInMean = nullptr;
            S.ensure (ctx, idx (Memory::D_IN_MEAN), 8u * sizeof (cl_float), false);
            if (detail::stage_out (staging)) { S.ensure (ctx, idx (Memory::H_OUT_T_K), 8u * sizeof (cl_float), true); hPtrOutTk = S.host<cl_float> (idx (Memory::H_OUT_T_K)); } else hPtrOutTk = nullptr;
            S.ensure (ctx, idx (Memory::D_OUT_T_K), 8u * sizeof (cl_float), false);
        }
        void write (Memory mem = Memory::D_IN_S, void *ptr = nullptr, bool block = CL_FALSE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_in (staging)) return;
            switch (mem)
            {
                case Memory::D_IN_S:
                    if (ptr != nullptr) std::memcpy (hPtrInS, ptr, S.bytes[idx (Memory::D_IN_S)]);
                    queue.enqueueWriteBuffer (S.at (idx (Memory::D_IN_S)), block, 0, S.bytes[idx (Memory::D_IN_S)], hPtrInS);
                    break;
                case Memory::D_IN_MEAN:
                    if (ptr != nullptr) std::memcpy (hPtrInMean, ptr, S.bytes[idx (Memory::D_IN_MEAN)]);
                    queue.enqueueWriteBuffer (S.at (idx (Memory::D_IN_MEAN)), block, 0, S.bytes[idx (Memory::D_IN_MEAN)], hPtrInMean);
                    break;
                default: break;
            }
        }
        void* read (Memory mem = Memory::H_OUT_T_K, bool block = CL_TRUE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_out (staging)) return nullptr;
            switch (mem)
            {
                case Memory::H_OUT_T_K:
                    queue.enqueueReadBuffer (S.at (idx (Memory::D_OUT_T_K)), block, 0, S.bytes[idx (Memory::D_OUT_T_K)], hPtrOutTk);
                    return hPtrOutTk;
                default: return nullptr;
            }
        }
        void run (const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            cl::check (icp_power_method (ctx (), (const float *) S.dev (idx (Memory::D_IN_S)), (const float *) S.dev (idx (Memory::D_IN_MEAN)), (float *) S.dev (idx (Memory::D_OUT_T_K))));
        }
        /*! \brief Timed execution (ms), the counterpart of run (clutils::GPUTimer&). */
        template <typename period>
        double run (clutils::GPUTimer<period> &timer, const std::vector<cl::Event> *events = nullptr)
        {
            timer.start (); run (events); return timer.stop ();
        }
        cl_float *hPtrInS = nullptr;  /*!< mapping of the H_IN_S staging buffer */
        cl_float *hPtrInMean = nullptr;  /*!< mapping of the H_IN_MEAN staging buffer */
        cl_float *hPtrOutTk = nullptr;  /*!< mapping of the H_OUT_T_K staging buffer */
    private:
        static unsigned idx (Memory m) { return static_cast<unsigned> (m); }
        clutils::CLEnv &env;
        clutils::CLEnvInfo<1> info;
        cl::Context ctx;
        cl::CommandQueue queue;
        Staging staging = Staging::IO;
        detail::Slots<6> S;
    };
    /*! \brief ICPMean<REGULAR|WEIGHTED> stage (algorithms.hpp:624-837). */
    template <ICPMeanConfig C>
    class ICPMean
    {
    public:
        enum class Memory : uint8_t { H_IN_F, H_IN_M, H_IN_W, H_IN_SUM_W, H_OUT, D_IN_F, D_IN_M, D_IN_W, D_IN_SUM_W, D_GM, D_OUT };
        ICPMean (clutils::CLEnv &_env, clutils::CLEnvInfo<1> _info) : env (_env), info (_info), ctx (_env.getContext (_info.ctxIdx)), queue (ctx) {}
        cl::Memory& get (Memory mem) { return S.at (idx (mem)); }
        void init (unsigned int _n, Staging _staging = Staging::IO)
        {
            n = _n; staging = _staging;
            const size_t fm = (size_t) n * 8u * sizeof (cl_float);
            const bool w = (C == ICPMeanConfig::WEIGHTED);
            if (detail::stage_in (staging))
            {
                S.ensure (ctx, idx (Memory::H_IN_F), fm, true); hPtrInF = S.template host<cl_float> (idx (Memory::H_IN_F));
                S.ensure (ctx, idx (Memory::H_IN_M), fm, true); hPtrInM = S.template host<cl_float> (idx (Memory::H_IN_M));
                if (w)
                {
                    S.ensure (ctx, idx (Memory::H_IN_W), (size_t) n * sizeof (cl_float), true); hPtrInW = S.template host<cl_float> (idx (Memory::H_IN_W));
                    S.ensure (ctx, idx (Memory::H_IN_SUM_W), sizeof (cl_double), true); hPtrInSW = S.template host<cl_double> (idx (Memory::H_IN_SUM_W));
                }
            }
            else { hPtrInF = hPtrInM = hPtrInW = nullptr; hPtrInSW = nullptr; }
            if (detail::stage_out (staging)) { S.ensure (ctx, idx (Memory::H_OUT), 8u * sizeof (cl_float), true); hPtrOut = S.template host<cl_float> (idx (Memory::H_OUT)); }
            else hPtrOut = nullptr;
            S.ensure (ctx, idx (Memory::D_IN_F), fm, false);
            S.ensure (ctx, idx (Memory::D_IN_M), fm, false);
            if (w) { S.ensure (ctx, idx (Memory::D_IN_W), (size_t) n * sizeof (cl_float), false); S.ensure (ctx, idx (Memory::D_IN_SUM_W), sizeof (cl_double), false); }
            S.ensure (ctx, idx (Memory::D_OUT), 8u * sizeof (cl_float), false);
        }
        void write (Memory mem = Memory::D_IN_F, void *ptr = nullptr, bool block = CL_FALSE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_in (staging)) return;
            void *h = nullptr;
            switch (mem)
            {
                case Memory::D_IN_F: h = hPtrInF; break;
                case Memory::D_IN_M: h = hPtrInM; break;
                case Memory::D_IN_W: h = hPtrInW; break;
                case Memory::D_IN_SUM_W: h = hPtrInSW; break;
                default: return;
            }
            if (h == nullptr) return;
            if (ptr != nullptr) std::memcpy (h, ptr, S.bytes[idx (mem)]);
            queue.enqueueWriteBuffer (S.at (idx (mem)), block, 0, S.bytes[idx (mem)], h);
        }
        void* read (Memory mem = Memory::H_OUT, bool block = CL_TRUE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_out (staging) || mem != Memory::H_OUT) return nullptr;
            queue.enqueueReadBuffer (S.at (idx (Memory::D_OUT)), block, 0, 8u * sizeof (cl_float), hPtrOut);
            return hPtrOut;
        }
        void run (const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (C == ICPMeanConfig::WEIGHTED)
                cl::check (icp_mean_weighted (ctx (), (const float *) S.dev (idx (Memory::D_IN_F)), (const float *) S.dev (idx (Memory::D_IN_M)),
                                              (const float *) S.dev (idx (Memory::D_IN_W)), (const double *) S.dev (idx (Memory::D_IN_SUM_W)),
                                              (float *) S.dev (idx (Memory::D_OUT)), n));
            else
                cl::check (icp_mean (ctx (), (const float *) S.dev (idx (Memory::D_IN_F)), (const float *) S.dev (idx (Memory::D_IN_M)),
                                     (float *) S.dev (idx (Memory::D_OUT)), n));
        }
        template <typename period>
        double run (clutils::GPUTimer<period> &timer, const std::vector<cl::Event> *events = nullptr)
        { timer.start (); run (events); return timer.stop (); }
        cl_float *hPtrInF = nullptr, *hPtrInM = nullptr, *hPtrInW = nullptr, *hPtrOut = nullptr;
        cl_double *hPtrInSW = nullptr;
    private:
        static unsigned idx (Memory m) { return static_cast<unsigned> (m); }
        clutils::CLEnv &env;
        clutils::CLEnvInfo<1> info;
        cl::Context ctx;
        cl::CommandQueue queue;
        Staging staging = Staging::IO;
        unsigned int n = 0;
        detail::Slots<11> S;
    };

    /*! \brief ICPS<REGULAR|WEIGHTED> stage (algorithms.hpp:990-1185). */
    template <ICPSConfig C>
    class ICPS
    {
    public:
        enum class Memory : uint8_t { H_IN_DEV_M, H_IN_DEV_F, H_IN_W, H_OUT, D_IN_DEV_M, D_IN_DEV_F, D_IN_W, D_SIJ, D_OUT };
        ICPS (clutils::CLEnv &_env, clutils::CLEnvInfo<1> _info) : env (_env), info (_info), ctx (_env.getContext (_info.ctxIdx)), queue (ctx) {}
        cl::Memory& get (Memory mem) { return S.at (idx (mem)); }
        void init (unsigned int _m, float _c = 1e-6f, Staging _staging = Staging::IO)
        {
            m = _m; c = _c; staging = _staging;
            const size_t dv = (size_t) m * 4u * sizeof (cl_float);
            const bool w = (C == ICPSConfig::WEIGHTED);
            if (detail::stage_in (staging))
            {
                S.ensure (ctx, idx (Memory::H_IN_DEV_M), dv, true); hPtrInDevM = S.template host<cl_float> (idx (Memory::H_IN_DEV_M));
                S.ensure (ctx, idx (Memory::H_IN_DEV_F), dv, true); hPtrInDevF = S.template host<cl_float> (idx (Memory::H_IN_DEV_F));
                if (w) { S.ensure (ctx, idx (Memory::H_IN_W), (size_t) m * sizeof (cl_float), true); hPtrInW = S.template host<cl_float> (idx (Memory::H_IN_W)); }
            }
            else hPtrInDevM = hPtrInDevF = hPtrInW = nullptr;
            if (detail::stage_out (staging)) { S.ensure (ctx, idx (Memory::H_OUT), 11u * sizeof (cl_float), true); hPtrOut = S.template host<cl_float> (idx (Memory::H_OUT)); }
            else hPtrOut = nullptr;
            S.ensure (ctx, idx (Memory::D_IN_DEV_M), dv, false);
            S.ensure (ctx, idx (Memory::D_IN_DEV_F), dv, false);
            if (w) S.ensure (ctx, idx (Memory::D_IN_W), (size_t) m * sizeof (cl_float), false);
            S.ensure (ctx, idx (Memory::D_OUT), 11u * sizeof (cl_float), false);
        }
        void write (Memory mem = Memory::D_IN_DEV_M, void *ptr = nullptr, bool block = CL_FALSE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_in (staging)) return;
            void *h = nullptr;
            switch (mem)
            {
                case Memory::D_IN_DEV_M: h = hPtrInDevM; break;
                case Memory::D_IN_DEV_F: h = hPtrInDevF; break;
                case Memory::D_IN_W: h = hPtrInW; break;
                default: return;
            }
            if (h == nullptr) return;
            if (ptr != nullptr) std::memcpy (h, ptr, S.bytes[idx (mem)]);
            queue.enqueueWriteBuffer (S.at (idx (mem)), block, 0, S.bytes[idx (mem)], h);
        }
        void* read (Memory mem = Memory::H_OUT, bool block = CL_TRUE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_out (staging) || mem != Memory::H_OUT) return nullptr;
            queue.enqueueReadBuffer (S.at (idx (Memory::D_OUT)), block, 0, 11u * sizeof (cl_float), hPtrOut);
            return hPtrOut;
        }
        void run (const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            cl::check (icp_sij (ctx (), (const float *) S.dev (idx (Memory::D_IN_DEV_M)), (const float *) S.dev (idx (Memory::D_IN_DEV_F)),
                                C == ICPSConfig::WEIGHTED ? (const float *) S.dev (idx (Memory::D_IN_W)) : nullptr,
                                (float *) S.dev (idx (Memory::D_OUT)), m, c));
        }
        template <typename period>
        double run (clutils::GPUTimer<period> &timer, const std::vector<cl::Event> *events = nullptr)
        { timer.start (); run (events); return timer.stop (); }
        float getScaling () { return c; }
        void setScaling (float _c) { c = _c; }
        cl_float *hPtrInDevM = nullptr, *hPtrInDevF = nullptr, *hPtrInW = nullptr, *hPtrOut = nullptr;
    private:
        static unsigned idx (Memory m_) { return static_cast<unsigned> (m_); }
        clutils::CLEnv &env;
        clutils::CLEnvInfo<1> info;
        cl::Context ctx;
        cl::CommandQueue queue;
        Staging staging = Staging::IO;
        unsigned int m = 0;
        float c = 1e-6f;
        detail::Slots<9> S;
    };

    /*! \brief ICPTransform<QUATERNION|MATRIX> stage (algorithms.hpp:1239-1424). */
    template <ICPTransformConfig C>
    class ICPTransform
    {
    public:
        enum class Memory : uint8_t { H_IN_M, H_IN_T, H_OUT, D_IN_M, D_IN_T, D_OUT };
        ICPTransform (clutils::CLEnv &_env, clutils::CLEnvInfo<1> _info) : env (_env), info (_info), ctx (_env.getContext (_info.ctxIdx)), queue (ctx) {}
        cl::Memory& get (Memory mem) { return S.at (idx (mem)); }
        void init (unsigned int _m, Staging _staging = Staging::IO)
        {
            m = _m; staging = _staging;
            const size_t pm = (size_t) m * 8u * sizeof (cl_float);
            const size_t pt = (C == ICPTransformConfig::QUATERNION ? 8u : 16u) * sizeof (cl_float);
            if (detail::stage_in (staging))
            {
                S.ensure (ctx, idx (Memory::H_IN_M), pm, true); hPtrInM = S.template host<cl_float> (idx (Memory::H_IN_M));
                S.ensure (ctx, idx (Memory::H_IN_T), pt, true); hPtrInT = S.template host<cl_float> (idx (Memory::H_IN_T));
            }
            else hPtrInM = hPtrInT = nullptr;
            if (detail::stage_out (staging)) { S.ensure (ctx, idx (Memory::H_OUT), pm, true); hPtrOut = S.template host<cl_float> (idx (Memory::H_OUT)); }
            else hPtrOut = nullptr;
            S.ensure (ctx, idx (Memory::D_IN_M), pm, false);
            S.ensure (ctx, idx (Memory::D_IN_T), pt, false);
            S.ensure (ctx, idx (Memory::D_OUT), pm, false);
        }
        void write (Memory mem = Memory::D_IN_M, void *ptr = nullptr, bool block = CL_FALSE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_in (staging)) return;
            void *h = (mem == Memory::D_IN_M) ? (void *) hPtrInM : (mem == Memory::D_IN_T) ? (void *) hPtrInT : nullptr;
            if (h == nullptr) return;
            if (ptr != nullptr) std::memcpy (h, ptr, S.bytes[idx (mem)]);
            queue.enqueueWriteBuffer (S.at (idx (mem)), block, 0, S.bytes[idx (mem)], h);
        }
        void* read (Memory mem = Memory::H_OUT, bool block = CL_TRUE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_out (staging) || mem != Memory::H_OUT) return nullptr;
            queue.enqueueReadBuffer (S.at (idx (Memory::D_OUT)), block, 0, S.bytes[idx (Memory::D_OUT)], hPtrOut);
            return hPtrOut;
        }
        void run (const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (C == ICPTransformConfig::QUATERNION)
                cl::check (icp_transform_quaternion (ctx (), (const float *) S.dev (idx (Memory::D_IN_M)), (const float *) S.dev (idx (Memory::D_IN_T)),
                                                     (float *) S.dev (idx (Memory::D_OUT)), m));
            else
                cl::check (icp_transform_matrix (ctx (), (const float *) S.dev (idx (Memory::D_IN_M)), (const float *) S.dev (idx (Memory::D_IN_T)),
                                                 (float *) S.dev (idx (Memory::D_OUT)), m));
        }
        template <typename period>
        double run (clutils::GPUTimer<period> &timer, const std::vector<cl::Event> *events = nullptr)
        { timer.start (); run (events); return timer.stop (); }
        cl_float *hPtrInM = nullptr, *hPtrInT = nullptr, *hPtrOut = nullptr;
    private:
        static unsigned idx (Memory m_) { return static_cast<unsigned> (m_); }
        clutils::CLEnv &env;
        clutils::CLEnvInfo<1> info;
        cl::Context ctx;
        cl::CommandQueue queue;
        Staging staging = Staging::IO;
        unsigned int m = 0;
        detail::Slots<6> S;
    };

    /*! \brief Reduce<MIN,float> / Reduce<MAX,uint> / Reduce<SUM,float> (algorithms.hpp:83-185). */
    template <ReduceConfig C, typename T>
    class Reduce
    {
    public:
        enum class Memory : uint8_t { H_IN, H_OUT, D_IN, D_RED, D_OUT };
        Reduce (clutils::CLEnv &_env, clutils::CLEnvInfo<1> _info) : env (_env), info (_info), ctx (_env.getContext (_info.ctxIdx)), queue (ctx) {}
        cl::Memory& get (Memory mem) { return S.at (idx (mem)); }
        void init (unsigned int _cols, unsigned int _rows, Staging _staging = Staging::IO)
        {
            cols = _cols; rows = _rows; staging = _staging;
            const size_t in = (size_t) cols * rows * sizeof (T), out = (size_t) rows * sizeof (T);
            if (detail::stage_in (staging)) { S.ensure (ctx, idx (Memory::H_IN), in, true); hPtrIn = S.template host<T> (idx (Memory::H_IN)); } else hPtrIn = nullptr;
            if (detail::stage_out (staging)) { S.ensure (ctx, idx (Memory::H_OUT), out, true); hPtrOut = S.template host<T> (idx (Memory::H_OUT)); } else hPtrOut = nullptr;
            S.ensure (ctx, idx (Memory::D_IN), in, false);
            S.ensure (ctx, idx (Memory::D_OUT), out, false);
        }
        void write (Memory mem = Memory::D_IN, void *ptr = nullptr, bool block = CL_FALSE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_in (staging) || mem != Memory::D_IN) return;
            if (ptr != nullptr) std::memcpy (hPtrIn, ptr, S.bytes[idx (Memory::D_IN)]);
            queue.enqueueWriteBuffer (S.at (idx (Memory::D_IN)), block, 0, S.bytes[idx (Memory::D_IN)], hPtrIn);
        }
        void* read (Memory mem = Memory::H_OUT, bool block = CL_TRUE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_out (staging) || mem != Memory::H_OUT) return nullptr;
            queue.enqueueReadBuffer (S.at (idx (Memory::D_OUT)), block, 0, S.bytes[idx (Memory::D_OUT)], hPtrOut);
            return hPtrOut;
        }
        void run (const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (C == ReduceConfig::MIN) cl::check (icp_reduce_min_f (ctx (), (const float *) S.dev (idx (Memory::D_IN)), cols, rows, (float *) S.dev (idx (Memory::D_OUT))));
            else if (C == ReduceConfig::MAX) cl::check (icp_reduce_max_ui (ctx (), (const uint32_t *) S.dev (idx (Memory::D_IN)), cols, rows, (uint32_t *) S.dev (idx (Memory::D_OUT))));
            else cl::check (icp_reduce_sum_f (ctx (), (const float *) S.dev (idx (Memory::D_IN)), cols, rows, (float *) S.dev (idx (Memory::D_OUT))));
        }
        template <typename period>
        double run (clutils::GPUTimer<period> &timer, const std::vector<cl::Event> *events = nullptr)
        { timer.start (); run (events); return timer.stop (); }
        T *hPtrIn = nullptr, *hPtrOut = nullptr;
    private:
        static unsigned idx (Memory m_) { return static_cast<unsigned> (m_); }
        clutils::CLEnv &env;
        clutils::CLEnvInfo<1> info;
        cl::Context ctx;
        cl::CommandQueue queue;
        Staging staging = Staging::IO;
        unsigned int cols = 0, rows = 0;
        detail::Slots<5> S;
    };

    /*! \brief Scan<INCLUSIVE|EXCLUSIVE,int> (algorithms.hpp:207-289). */
    template <ScanConfig C, typename T>
    class Scan
    {
    public:
        enum class Memory : uint8_t { H_IN, H_OUT, D_IN, D_SUMS, D_OUT };
        Scan (clutils::CLEnv &_env, clutils::CLEnvInfo<1> _info) : env (_env), info (_info), ctx (_env.getContext (_info.ctxIdx)), queue (ctx) {}
        cl::Memory& get (Memory mem) { return S.at (idx (mem)); }
        void init (unsigned int _cols, unsigned int _rows, Staging _staging = Staging::IO)
        {
            cols = _cols; rows = _rows; staging = _staging;
            const size_t sz = (size_t) cols * rows * sizeof (T);
            if (detail::stage_in (staging)) { S.ensure (ctx, idx (Memory::H_IN), sz, true); hPtrIn = S.template host<T> (idx (Memory::H_IN)); } else hPtrIn = nullptr;
            if (detail::stage_out (staging)) { S.ensure (ctx, idx (Memory::H_OUT), sz, true); hPtrOut = S.template host<T> (idx (Memory::H_OUT)); } else hPtrOut = nullptr;
            S.ensure (ctx, idx (Memory::D_IN), sz, false);
            S.ensure (ctx, idx (Memory::D_OUT), sz, false);
        }
        void write (Memory mem = Memory::D_IN, void *ptr = nullptr, bool block = CL_FALSE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_in (staging) || mem != Memory::D_IN) return;
            if (ptr != nullptr) std::memcpy (hPtrIn, ptr, S.bytes[idx (Memory::D_IN)]);
            queue.enqueueWriteBuffer (S.at (idx (Memory::D_IN)), block, 0, S.bytes[idx (Memory::D_IN)], hPtrIn);
        }
        void* read (Memory mem = Memory::H_OUT, bool block = CL_TRUE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_out (staging) || mem != Memory::H_OUT) return nullptr;
            queue.enqueueReadBuffer (S.at (idx (Memory::D_OUT)), block, 0, S.bytes[idx (Memory::D_OUT)], hPtrOut);
            return hPtrOut;
        }
        void run (const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            cl::check (icp_scan_i (ctx (), (const int32_t *) S.dev (idx (Memory::D_IN)), cols, rows, C == ScanConfig::INCLUSIVE ? 1 : 0,
                                   (int32_t *) S.dev (idx (Memory::D_OUT))));
        }
        template <typename period>
        double run (clutils::GPUTimer<period> &timer, const std::vector<cl::Event> *events = nullptr)
        { timer.start (); run (events); return timer.stop (); }
        T *hPtrIn = nullptr, *hPtrOut = nullptr;
    private:
        static unsigned idx (Memory m_) { return static_cast<unsigned> (m_); }
        clutils::CLEnv &env;
        clutils::CLEnvInfo<1> info;
        cl::Context ctx;
        cl::CommandQueue queue;
        Staging staging = Staging::IO;
        unsigned int cols = 0, rows = 0;
        detail::Slots<5> S;
    };

    /*! \brief ICPStep<CR,CW>: one ICP iteration per run () (algorithms.hpp:1613-2401).  The stage objects of the
     *         reference (fReps, rbcC, transform, rbcS, weights, means, devs, matrixS, powMethod) are replaced by the
     *         device-resident engine behind icp_step_*; buffer wiring (algorithms.cpp:4495-4581) happens inside it. */
    template <ICPStepConfigT CR, ICPStepConfigW CW>
    class ICPStep
    {
    public:
        enum class Memory : uint8_t { H_IN_F, H_IN_M, H_IO_T, D_IN_F, D_IN_M, D_IO_T };
        ICPStep (clutils::CLEnv &_env, clutils::CLEnvInfo<1> _infoRBC, clutils::CLEnvInfo<1> _infoICP)
            : env (_env), infoRBC (_infoRBC), infoICP (_infoICP), context (_env.getContext (_infoICP.ctxIdx)), queue (context)
        {
            static_assert (CR != ICPStepConfigT::JACOBI, "ICPStepConfigT::JACOBI is a todo in the reference as well");
            icp_step *p = nullptr;
            cl::check (icp_step_create (context (), CR == ICPStepConfigT::POWER_METHOD ? ICP_ROT_POWER_METHOD : ICP_ROT_EIGEN,
                                        CW == ICPStepConfigW::WEIGHTED ? ICP_W_WEIGHTED : ICP_W_REGULAR, &p));
            h.reset (p, [] (icp_step *q_) { icp_step_destroy (q_); });
        }
        cl::Memory& get (Memory mem) { return S.at (idx (mem)); }
        void init (unsigned int _m, unsigned int _nr, float _a = 1e2f, float _c = 1e-6f, Staging _staging = Staging::IO)
        {
            m = _m; nr = _nr; a = _a; c = _c; staging = _staging;
            const size_t fm = (size_t) m * 8u * sizeof (cl_float), pt = 8u * sizeof (cl_float);
            if (detail::stage_in (staging))
            {
                S.ensure (context, idx (Memory::H_IN_F), fm, true); hPtrInF = S.template host<cl_float> (idx (Memory::H_IN_F));
                S.ensure (context, idx (Memory::H_IN_M), fm, true); hPtrInM = S.template host<cl_float> (idx (Memory::H_IN_M));
            }
            else hPtrInF = hPtrInM = nullptr;
            S.ensure (context, idx (Memory::H_IO_T), pt, true); hPtrIOT = S.template host<cl_float> (idx (Memory::H_IO_T));
            S.ensure (context, idx (Memory::D_IN_F), fm, false);
            S.ensure (context, idx (Memory::D_IN_M), fm, false);
            S.ensure (context, idx (Memory::D_IO_T), pt, false);
            cl::check (icp_step_bind (h.get (), ICP_MEM_D_IN_F, S.dev (idx (Memory::D_IN_F))));
            cl::check (icp_step_bind (h.get (), ICP_MEM_D_IN_M, S.dev (idx (Memory::D_IN_M))));
            cl::check (icp_step_bind (h.get (), ICP_MEM_D_IO_T, S.dev (idx (Memory::D_IO_T))));
            cl::check (icp_step_init (h.get (), m, nr, a, c, 0, 0));      // loads the identity pose, R = I, t = 0, s = 1
            refresh ();
        }
        void write (Memory mem = Memory::D_IN_F, void *ptr = nullptr, bool block = CL_FALSE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_in (staging)) return;
            switch (mem)
            {
                case Memory::D_IN_F:
                    if (ptr != nullptr) std::memcpy (hPtrInF, ptr, S.bytes[idx (Memory::D_IN_F)]);
                    queue.enqueueWriteBuffer (S.at (idx (Memory::D_IN_F)), block, 0, S.bytes[idx (Memory::D_IN_F)], hPtrInF);
                    break;
                case Memory::D_IN_M:
                    if (ptr != nullptr) std::memcpy (hPtrInM, ptr, S.bytes[idx (Memory::D_IN_M)]);
                    queue.enqueueWriteBuffer (S.at (idx (Memory::D_IN_M)), block, 0, S.bytes[idx (Memory::D_IN_M)], hPtrInM);
                    break;
                case Memory::D_IO_T:
                    if (ptr != nullptr) std::memcpy (hPtrIOT, ptr, 8u * sizeof (cl_float));
                    queue.enqueueWriteBuffer (S.at (idx (Memory::D_IO_T)), block, 0, 8u * sizeof (cl_float), hPtrIOT);
                    break;
                default: break;
            }
        }
        void* read (Memory mem = Memory::H_IO_T, bool block = CL_TRUE,
                    const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            if (!detail::stage_out (staging) || mem != Memory::H_IO_T) return nullptr;
            queue.enqueueReadBuffer (S.at (idx (Memory::D_IO_T)), block, 0, 8u * sizeof (cl_float), hPtrIOT);
            return hPtrIOT;
        }
        void buildRBC (const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            (void) events; (void) event;
            cl::check (icp_step_build_rbc (h.get ()));
        }
        /*! One iteration.  Like the reference it ends with a blocking read of the incremental / accumulated pose
         *  (algorithms.cpp:4681), which refreshes the public members. */
        void run (const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr, bool config = false)
        {
            (void) events; (void) event; (void) config;     // no host-side search configuration is needed
            cl::check (icp_step_run (h.get (), 1));
            refresh ();
        }
        /*! n iterations back to back on the device (no host round trip); members refreshed at the end. */
        void run (unsigned int n_iterations)
        {
            cl::check (icp_step_run (h.get (), n_iterations));
            refresh ();
        }
        template <typename period>
        double run (clutils::GPUTimer<period> &timer, const std::vector<cl::Event> *events = nullptr, bool config = false)
        {
            (void) events; (void) config;
            timer.start (); cl::check (icp_step_run (h.get (), 1)); const double ms = timer.stop ();
            refresh ();
            return ms;
        }
        float getAlpha () { return a; }
        void setAlpha (float _a) { a = _a; cl::check (icp_step_set_alpha (h.get (), a)); }
        float getScaling () { return c; }
        void setScaling (float _c) { c = _c; cl::check (icp_step_set_scaling (h.get (), c)); }
        /*! Execution mode of the engine: ICP_MODE_FUSED (default) or ICP_MODE_STAGED (one kernel per reference stage). */
        void setMode (int mode) { cl::check (icp_step_set_mode (h.get (), mode)); }
        icp_step* handle () { return h.get (); }

        cl_float *hPtrInF = nullptr;  /*!< mapping of the input staging buffer for the fixed set */
        cl_float *hPtrInM = nullptr;  /*!< mapping of the input staging buffer for the moving set */
        cl_float *hPtrIOT = nullptr;  /*!< mapping of the input-output staging buffer for the pose {q, t, s} */
        Matrix3f Rk; Quaternionf qk; Vector3f tk; float sk = 1.f;   /*!< incremental estimate of iteration k */
        Matrix3f R; Quaternionf q; Vector3f t; float s = 1.f;        /*!< accumulated estimate up to iteration k */

    protected:
        void refresh ()
        {
            icp_state st;
            cl::check (icp_step_get_state (h.get (), &st));
            // through the public accessors only: works with the look-alikes of eigen_shim.hpp and with the real Eigen types
            for (int r_ = 0; r_ < 3; ++r_)
                for (int c_ = 0; c_ < 3; ++c_) { Rk (r_, c_) = st.Rk[r_ * 3 + c_]; R (r_, c_) = st.R[r_ * 3 + c_]; }
            qk = Quaternionf (st.qk[3], st.qk[0], st.qk[1], st.qk[2]);       // (w, x, y, z)
            q = Quaternionf (st.q[3], st.q[0], st.q[1], st.q[2]);
            for (int i_ = 0; i_ < 3; ++i_) { tk[i_] = st.tk[i_]; t[i_] = st.t[i_]; }
            sk = st.sk; s = st.s;
            k_dev = st.k;
        }
        static unsigned idx (Memory m_) { return static_cast<unsigned> (m_); }
        clutils::CLEnv &env;
        clutils::CLEnvInfo<1> infoRBC, infoICP;
        cl::Context context;
        cl::CommandQueue queue;
        Staging staging = Staging::IO;
        std::shared_ptr<icp_step> h;
        float a = 1e2f, c = 1e-6f;
        unsigned int m = 0, nr = 0, d = 8;
        unsigned int k_dev = 0;
        detail::Slots<6> S;
    };

    /*! \brief ICP<CR,CW>: the iterative registration driver (algorithms.hpp:2433-2496).  run () executes the whole
     *         `do { step } while (check ())` loop on the device (conditional CUDA graph). */
    template <ICPStepConfigT CR, ICPStepConfigW CW>
    class ICP : public ICPStep<CR, CW>
    {
    public:
        ICP (clutils::CLEnv &_env, clutils::CLEnvInfo<1> _infoRBC, clutils::CLEnvInfo<1> _infoICP) : ICPStep<CR, CW> (_env, _infoRBC, _infoICP) {}
        void init (unsigned int _m, unsigned int _nr, float _a = 1e2f, float _c = 1e-6f, unsigned int _max_iterations = 40,
                   double _angle_threshold = 0.001, double _translation_threshold = 0.01, Staging _staging = Staging::IO)
        {
            max_iterations = _max_iterations; angle_threshold = _angle_threshold; translation_threshold = _translation_threshold;
            ICPStep<CR, CW>::init (_m, _nr, _a, _c, _staging);
        }
        void buildRBC (const std::vector<cl::Event> *events = nullptr, cl::Event *event = nullptr)
        {
            ICPStep<CR, CW>::buildRBC (events, event);
            k = 0;
        }
        /*! Blocking, like the reference (queue.finish () at algorithms.cpp:4813). */
        void run ()
        {
            uint32_t kk = 0;
            cl::check (icp_run (this->h.get (), max_iterations, angle_threshold, translation_threshold, &kk));
            k = kk;
            this->refresh ();
        }
        /*! Profiling driver: exactly 40 steps, no convergence test (algorithms.hpp:2482-2494); returns the total in ms. */
        template <typename period>
        double run (clutils::GPUTimer<period> &timer)
        {
            timer.start ();
            cl::check (icp_step_run (this->h.get (), 40));
            const double ms = timer.stop ();
            this->refresh ();
            k = this->k_dev;
            return ms;
        }
        unsigned int getMaxIterations () { return max_iterations; }
        void setMaxIterations (unsigned int v) { max_iterations = v; }
        double getAngleThreshold () { return angle_threshold; }
        void setAngleThreshold (double v) { angle_threshold = v; }
        double getTranslationThreshold () { return translation_threshold; }
        void setTranslationThreshold (double v) { translation_threshold = v; }
        unsigned int k = 0;   /*!< iterations performed by the last run () */
    protected:
        unsigned int max_iterations = 40;
        double angle_threshold = 0.001;
        double translation_threshold = 0.01;
    };

    /*! \brief ICPBatch<CR,CW>: many independent registrations per call, on one or several GPUs of the box (throughput
     *         mode; no counterpart in the reference, which registers one pair per ICP<CR,CW> object on one queue).
     *  \details Same shape as ICP<CR,CW> (algorithms.hpp:2433-2496): ctor (env, infoRBC, infoICP) . init . write . run .
     *           read.  Every pair goes through exactly the arithmetic of ICP<CR,CW>::buildRBC () + a fixed number of
     *           ICPStep::run () iterations (the profiling driver of algorithms.hpp:2482-2494), so pose p equals what
     *           ICP<CR,CW> computes for pair p.  Device d owns a contiguous block of pairs end to end (icp_multi_*): one
     *           host thread and one context / stream set per GPU inside run (), no collective, only poses come back. */
    template <ICPStepConfigT CR, ICPStepConfigW CW>
    class ICPBatch
    {
    public:
        enum class Memory : uint8_t { H_IN_F, H_IN_M, H_OUT_T };
        ICPBatch (clutils::CLEnv &_env, clutils::CLEnvInfo<1> _infoRBC, clutils::CLEnvInfo<1> _infoICP) : env (_env), infoRBC (_infoRBC), infoICP (_infoICP)
        {
            static_assert (CR != ICPStepConfigT::JACOBI, "ICPStepConfigT::JACOBI is a todo in the reference as well");
        }
        /*! \param n_devices 0 = every visible GPU.  Host staging buffers are pinned (CL_MEM_ALLOC_HOST_PTR in the reference). */
        void init (unsigned int _n_pairs, unsigned int _m, unsigned int _nr, float _a = 1e2f, float _c = 1e-6f,
                   unsigned int _iterations = 40, int n_devices = 0)
        {
            n_pairs = _n_pairs; m = _m; nr = _nr; a = _a; c = _c; iterations = _iterations;
            icp_multi *p = nullptr;
            cl::check (icp_multi_create (n_devices, nullptr, CR == ICPStepConfigT::POWER_METHOD ? ICP_ROT_POWER_METHOD : ICP_ROT_EIGEN,
                                         CW == ICPStepConfigW::WEIGHTED ? ICP_W_WEIGHTED : ICP_W_REGULAR, n_pairs, m, nr, a, c, 0, 0, &p));
            h.reset (p, [] (icp_multi *q_) { icp_multi_destroy (q_); });
            const size_t fm = (size_t) n_pairs * m * 8u * sizeof (cl_float);
            hostF = pinned (fm); hostM = pinned (fm); hostT = pinned ((size_t) n_pairs * 8u * sizeof (cl_float));
            hPtrInF = static_cast<cl_float *> (hostF.get ()); hPtrInM = static_cast<cl_float *> (hostM.get ());
            hPtrOutT = static_cast<cl_float *> (hostT.get ());
        }
        /*! Copies pair `pair` (m x 8 floats) into the staging buffer of the fixed (H_IN_F) or moving (H_IN_M) sets. */
        void write (Memory mem, unsigned int pair, const void *ptr)
        {
            cl_float *dst = (mem == Memory::H_IN_F) ? hPtrInF : hPtrInM;
            std::memcpy (dst + (size_t) pair * m * 8u, ptr, (size_t) m * 8u * sizeof (cl_float));
        }
        /*! buildRBC + `iterations` steps of every pair; blocking.  Poses {q, t, s} land in hPtrOutT [n_pairs][8]. */
        void run () { cl::check (icp_multi_register_host (h.get (), hPtrInF, hPtrInM, iterations, hPtrOutT)); }
        /*! Pose of pair `pair` ({qx qy qz qw tx ty tz s}, the D_IO_T layout of algorithms.hpp:2245-2254). */
        cl_float* read (unsigned int pair = 0) { return hPtrOutT + (size_t) pair * 8u; }
        int devices () { return icp_multi_devices (h.get ()); }
        unsigned int getIterations () { return iterations; }
        void setIterations (unsigned int v) { iterations = v; }

        cl_float *hPtrInF = nullptr;   /*!< [n_pairs][m][8] fixed sets */
        cl_float *hPtrInM = nullptr;   /*!< [n_pairs][m][8] moving sets */
        cl_float *hPtrOutT = nullptr;  /*!< [n_pairs][8] poses */
    private:
        static std::shared_ptr<void> pinned (size_t bytes)
        {
            void *p = nullptr;
            cl::check (icp_host_alloc (bytes, &p));
            return std::shared_ptr<void> (p, [] (void *q_) { icp_host_free (q_); });
        }
        clutils::CLEnv &env;
        clutils::CLEnvInfo<1> infoRBC, infoICP;
        std::shared_ptr<icp_multi> h;
        std::shared_ptr<void> hostF, hostM, hostT;
        unsigned int n_pairs = 0, m = 0, nr = 0, iterations = 40;
        float a = 1e2f, c = 1e-6f;
    };

}
}

#endif  // ICP_ALGORITHMS_HPP
