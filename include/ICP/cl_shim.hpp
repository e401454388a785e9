/*! \file cl_shim.hpp
 *  \brief Minimal stand-ins for the OpenCL C++ binding types (cl::*) and the CLUtils types (clutils::*) that appear
 *         in the reference's algorithm-class API (/root/reference/include/ICP/algorithms.hpp:35-39), implemented on
 *         top of the C ABI of libicp_b200.so.  Only what the class API needs: a ref-counted device buffer handle,
 *         a context/queue pair (= one icp_ctx: device + in-order CUDA stream), an event placeholder and a GPU timer.
 *
 *  No OpenCL is involved.  `cl::Buffer` owns CUDA device memory (icp_malloc / icp_free); copying a handle shares the
 *  allocation exactly like copying a cl::Buffer shares the cl_mem (reference ownership rules).
 */
#ifndef ICP_CL_SHIM_HPP
#define ICP_CL_SHIM_HPP

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <iostream>
#include <memory>
#include <ratio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../icp_b200.h"

#ifndef CL_TRUE
#define CL_TRUE 1
#define CL_FALSE 0
#endif

typedef float cl_float;
typedef double cl_double;
typedef uint32_t cl_uint;
typedef int32_t cl_int;
typedef icp_dist_id rbc_dist_id;                     /* RBC/data_types.hpp */
typedef struct { cl_float s[4]; } cl_float4;
typedef struct { cl_float s[8]; } cl_float8;

enum { CL_MEM_READ_WRITE = 1, CL_MEM_WRITE_ONLY = 2, CL_MEM_READ_ONLY = 4, CL_MEM_ALLOC_HOST_PTR = 16 };
enum { CL_QUEUE_PROFILING_ENABLE = 2 };

namespace cl
{
    /*! \brief Thrown where the reference lets cl::Error propagate (runtime failures). */
    class Error : public std::runtime_error
    {
    public:
        Error (int code, const std::string &what_) : std::runtime_error (what_), code_ (code) {}
        int err () const { return code_; }
    private:
        int code_;
    };

    inline void check (int rc)
    {
        if (rc == ICP_OK) return;
        if (rc == ICP_ERR_CONFIG)
        {
            // reference behaviour for configuration errors: message on stderr, then exit (algorithms.cpp:164-168)
            std::cerr << icp_last_error () << std::endl;
            std::exit (EXIT_FAILURE);
        }
        throw Error (rc, icp_last_error ());
    }

    /*! \brief One device + one in-order stream (the reference uses one context and one in-order queue). */
    class Context
    {
    public:
        Context () {}
        explicit Context (int device)
        {
            icp_ctx *c = nullptr;
            check (icp_ctx_create (device, nullptr, &c));
            h.reset (c, [] (icp_ctx *p) { icp_ctx_destroy (p); });
        }
        icp_ctx* operator() () const { return h.get (); }
    private:
        std::shared_ptr<icp_ctx> h;
    };

    class Event {};

    /*! \brief Ref-counted memory handle; `mem () == nullptr` tests for "not created yet" like the reference does. */
    class Memory
    {
    public:
        Memory () {}
        void* operator() () const { return a ? a->ptr : nullptr; }
        size_t size () const { return a ? a->bytes : 0; }
    protected:
        struct Alloc
        {
            Context ctx; void *ptr; size_t bytes; bool host;
            ~Alloc () { if (ptr) { if (host) icp_host_free (ptr); else icp_free (ctx (), ptr); } }
        };
        std::shared_ptr<Alloc> a;
    };

    class Buffer : public Memory
    {
    public:
        Buffer () {}
        /*! CL_MEM_ALLOC_HOST_PTR => pinned host staging buffer, otherwise device memory. */
        Buffer (const Context &ctx, int flags, size_t bytes)
        {
            a = std::make_shared<Alloc> ();
            a->ctx = ctx; a->ptr = nullptr; a->bytes = bytes; a->host = (flags & CL_MEM_ALLOC_HOST_PTR) != 0;
            if (a->host) check (icp_host_alloc (bytes, &a->ptr));
            else check (icp_malloc (ctx (), bytes, &a->ptr));
        }
    };

    /*! \brief The in-order queue of a context. */
    class CommandQueue
    {
    public:
        CommandQueue () {}
        explicit CommandQueue (const Context &c) : ctx (c) {}
        void finish () const { check (icp_ctx_sync (ctx ())); }
        void flush () const {}
        /*! `offset` is a byte offset into the buffer, as in clEnqueueWriteBuffer / clEnqueueReadBuffer. */
        void enqueueWriteBuffer (const Memory &dst, bool block, size_t offset, size_t bytes, const void *src,
                                 const std::vector<Event>* = nullptr, Event* = nullptr) const
        {
            if (offset + bytes > dst.size ()) throw Error (ICP_ERR_ARG, "enqueueWriteBuffer: offset + size exceeds the buffer (CL_INVALID_VALUE)");
            check (icp_memcpy_h2d (ctx (), static_cast<char *> (dst ()) + offset, src, bytes, block ? 1 : 0));
        }
        void enqueueReadBuffer (const Memory &src, bool block, size_t offset, size_t bytes, void *dst,
                                const std::vector<Event>* = nullptr, Event* = nullptr) const
        {
            if (offset + bytes > src.size ()) throw Error (ICP_ERR_ARG, "enqueueReadBuffer: offset + size exceeds the buffer (CL_INVALID_VALUE)");
            check (icp_memcpy_d2h (ctx (), dst, static_cast<const char *> (src ()) + offset, bytes, block ? 1 : 0));
        }
        const Context& context () const { return ctx; }
    private:
        Context ctx;
    };
}

namespace clutils
{
    /*! \brief Replacement of clutils::CLEnvInfo<N>: indices of platform / device / context / queues / program. */
    template <unsigned N>
    struct CLEnvInfo
    {
        CLEnvInfo (unsigned p = 0, unsigned d = 0, unsigned c = 0, std::vector<unsigned> q = { 0 }, unsigned pg = 0)
            : pIdx (p), dIdx (d), ctxIdx (c), qIdx (q), pgIdx (pg) {}
        unsigned pIdx, dIdx, ctxIdx;
        std::vector<unsigned> qIdx;
        unsigned pgIdx;
    };

    /*! \brief Replacement of clutils::CLEnv: a list of contexts (one CUDA device + stream each). Programs are no-ops:
     *         the kernels are compiled into libicp_b200.so. */
    class CLEnv
    {
    public:
        CLEnv (const std::string & = std::string ()) {}
        cl::Context& addContext (unsigned device = 0) { contexts.emplace_back ((int) device); return contexts.back (); }
        cl::CommandQueue& addQueue (unsigned ctxIdx, unsigned = 0, int = 0) { queues.emplace_back (getContext (ctxIdx)); return queues.back (); }
        void addProgram (unsigned, const std::string &) {}
        void addProgram (unsigned, const std::vector<std::string> &) {}
        cl::Context& getContext (unsigned i = 0)
        {
            if (contexts.empty ()) addContext (0);
            return contexts.at (i < contexts.size () ? i : 0);
        }
        cl::CommandQueue getQueue (unsigned ctxIdx = 0, unsigned = 0) { return cl::CommandQueue (getContext (ctxIdx)); }
    private:
        // deques: addContext / addQueue hand out references that must survive later additions
        std::deque<cl::Context> contexts;
        std::deque<cl::CommandQueue> queues;
    };

    /*! \brief CUDA-event timer with the interface of clutils::GPUTimer (milliseconds). */
    template <typename period = std::milli>
    class GPUTimer
    {
    public:
        explicit GPUTimer (cl::Context &c) : ctx (c), last (0.0) {}
        void start () { cl::check (icp_timer_start (ctx ())); }
        double stop () { float ms = 0.f; cl::check (icp_timer_stop (ctx (), &ms)); last = ms; return last; }
        double duration () const { return last; }
        cl::Event& event () { return ev; }
        void wait () {}
    private:
        cl::Context ctx;
        cl::Event ev;
        double last;
    };
}

#endif  // ICP_CL_SHIM_HPP
