/*! \file registration.cpp
 *  \brief Headless frame-to-frame registration of two pc8d clouds with the drop-in classes of include/ICP.
 *  \details The counterpart of the reference's `icp_registration` app (examples/registration.cpp +
 *           src/ocl_icp_reg.cpp, ICPReg<RC,WC>::registerPC at :165-207) without the GLUT/OpenGL viewer:
 *           two 640x480 pc8d frames (raw little-endian f32, 640*480*8 values, data/README.md) are read from disk,
 *           128x128 landmarks are sampled from each (ICPLMs), the moving set is registered to the fixed set
 *           (ICP<RC,WC>: buildRBC + run), the final pose is applied to the full moving cloud (ICPTransform) and
 *           the summary the reference prints (iterations, latency, angle, axis, translation, scale) goes to stdout.
 *
 *  usage: registration <fixed.bin> <moving.bin> [--svd] [--regular] [--m 16384 --r 256 --alpha 200 --c 1e-6]
 *                      [--max-iterations 40] [--angle-threshold 0.001] [--translation-threshold 0.01]
 *                      [--out registered.bin] [--pose pose.txt]
 */
#include <ICP/algorithms.hpp>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace cl_algo::ICP;

namespace
{
    const unsigned int kWidth = 640, kHeight = 480, kPoints = kWidth * kHeight, kDim = 8;

    struct Options
    {
        std::string fixed, moving, out, pose;
        bool svd = false, regular = false;
        unsigned int m = 16384, r = 256, max_iterations = 40;
        float alpha = 2e2f, c = 1e-6f;                               // src/ocl_icp_reg.cpp:88
        double angle_threshold = 0.001, translation_threshold = 0.01;
    };

    bool load_pc8d (const std::string &path, std::vector<float> &cloud)
    {
        cloud.resize ((size_t) kPoints * kDim);
        FILE *f = std::fopen (path.c_str (), "rb");
        if (f == nullptr) { std::fprintf (stderr, "Error[registration]: cannot open %s\n", path.c_str ()); return false; }
        const size_t got = std::fread (cloud.data (), sizeof (float), cloud.size (), f);
        std::fclose (f);
        if (got != cloud.size ()) { std::fprintf (stderr, "Error[registration]: %s is not a 640x480 pc8d frame\n", path.c_str ()); return false; }
        return true;
    }

    bool parse (int argc, char **argv, Options &o)
    {
        std::vector<std::string> pos;
        for (int i = 1; i < argc; ++i)
        {
            const std::string a (argv[i]);
            auto next = [&] () -> const char * { return (i + 1 < argc) ? argv[++i] : "0"; };
            if (a == "--svd") o.svd = true;
            else if (a == "--regular") o.regular = true;
            else if (a == "--m") o.m = (unsigned) std::atoi (next ());
            else if (a == "--r") o.r = (unsigned) std::atoi (next ());
            else if (a == "--alpha") o.alpha = (float) std::atof (next ());
            else if (a == "--c") o.c = (float) std::atof (next ());
            else if (a == "--max-iterations") o.max_iterations = (unsigned) std::atoi (next ());
            else if (a == "--angle-threshold") o.angle_threshold = std::atof (next ());
            else if (a == "--translation-threshold") o.translation_threshold = std::atof (next ());
            else if (a == "--out") o.out = next ();
            else if (a == "--pose") o.pose = next ();
            else pos.push_back (a);
        }
        if (pos.size () != 2) return false;
        o.fixed = pos[0]; o.moving = pos[1];
        return true;
    }

    template <ICPStepConfigT RC, ICPStepConfigW WC>
    int registerPC (const Options &o, std::vector<float> &fixed, std::vector<float> &moving)
    {
        typedef ICPTransform<ICPTransformConfig::QUATERNION> Transform;
        clutils::CLEnv env;
        env.addContext (0);
        env.addQueue (0, 0);
        clutils::CLEnvInfo<1> infoRBC (0, 0, 0, { 0 }, 0), infoICP (0, 0, 0, { 0 }, 1);

        // landmark samplers feed the registration directly on the device (same wiring as ocl_icp_reg.cpp:103-121)
        ICPLMs fLM (env, infoICP), mLM (env, infoICP);
        fLM.init (Staging::I);
        mLM.init (Staging::I);
        ICP<RC, WC> reg (env, infoRBC, infoICP);
        reg.get (ICPStep<RC, WC>::Memory::D_IN_F) = fLM.get (ICPLMs::Memory::D_OUT);
        reg.get (ICPStep<RC, WC>::Memory::D_IN_M) = mLM.get (ICPLMs::Memory::D_OUT);
        reg.init (o.m, o.r, o.alpha, o.c, o.max_iterations, o.angle_threshold, o.translation_threshold, Staging::O);
        // the final pose is applied to the whole moving frame
        Transform transform (env, infoICP);
        transform.get (Transform::Memory::D_IN_M) = mLM.get (ICPLMs::Memory::D_IN);
        transform.get (Transform::Memory::D_IN_T) = reg.get (ICPStep<RC, WC>::Memory::D_IO_T);
        transform.init (kPoints, Staging::O);

        fLM.write (ICPLMs::Memory::D_IN, fixed.data ());
        mLM.write (ICPLMs::Memory::D_IN, moving.data (), CL_TRUE);
        fLM.run ();
        mLM.run ();

        reg.buildRBC ();
        const auto t0 = std::chrono::steady_clock::now ();
        reg.run ();                                                   // blocking, like the reference
        const double ms = std::chrono::duration<double, std::milli> (std::chrono::steady_clock::now () - t0).count ();
        transform.run ();
        const float *registered = (const float *) transform.read (Transform::Memory::H_OUT, CL_TRUE);

        const double sinth_2 = std::sqrt ((double) reg.q.x () * reg.q.x () + (double) reg.q.y () * reg.q.y () + (double) reg.q.z () * reg.q.z ());
        const double angle = 180.0 / M_PI * 2.0 * std::atan2 (sinth_2, (double) reg.q.w ());
        double axis[3] = { 0.0, 0.0, 0.0 };
        if (sinth_2 != 0.0) for (int i = 0; i < 3; ++i) axis[i] = reg.q.coeffs ()[i] / sinth_2;

        std::printf ("\n================\n\n");
        std::printf ("    Iterations            :    %u\n", reg.k);
        std::printf ("    Latency               :    %.4f ms\n", ms);
        std::printf ("    Rotation angle        :    %.6f degrees\n", angle);
        std::printf ("    Rotation axis         :    %.6f %.6f %.6f\n", axis[0], axis[1], axis[2]);
        std::printf ("    Translation vector    :    %.6f %.6f %.6f\n", reg.t[0], reg.t[1], reg.t[2]);
        std::printf ("    Scale                 :    %.6f\n", reg.s);

        if (!o.pose.empty ())
        {
            FILE *f = std::fopen (o.pose.c_str (), "w");
            if (f == nullptr) { std::fprintf (stderr, "Error[registration]: cannot write %s\n", o.pose.c_str ()); return EXIT_FAILURE; }
            std::fprintf (f, "%u\n", reg.k);
            std::fprintf (f, "%.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", reg.q.x (), reg.q.y (), reg.q.z (), reg.q.w (),
                          reg.t[0], reg.t[1], reg.t[2], reg.s);
            for (int i = 0; i < 3; ++i)
                std::fprintf (f, "%.9g %.9g %.9g %.9g\n", reg.s * reg.R (i, 0), reg.s * reg.R (i, 1), reg.s * reg.R (i, 2), reg.t[i]);
            std::fprintf (f, "0 0 0 1\n");
            std::fclose (f);
        }
        if (!o.out.empty ())
        {
            FILE *f = std::fopen (o.out.c_str (), "wb");
            if (f == nullptr || std::fwrite (registered, sizeof (float), (size_t) kPoints * kDim, f) != (size_t) kPoints * kDim)
            { std::fprintf (stderr, "Error[registration]: cannot write %s\n", o.out.c_str ()); return EXIT_FAILURE; }
            std::fclose (f);
        }
        return EXIT_SUCCESS;
    }
}

int main (int argc, char **argv)
{
    Options o;
    if (!parse (argc, argv, o))
    {
        std::fprintf (stderr, "usage: %s <fixed.bin> <moving.bin> [--svd] [--regular] [--m M --r R --alpha A --c C] [--max-iterations N]\n"
                              "          [--angle-threshold DEG] [--translation-threshold MM] [--out registered.bin] [--pose pose.txt]\n", argv[0]);
        return EXIT_FAILURE;
    }
    if (o.m != 16384) { std::fprintf (stderr, "Error[registration]: ICPLMs samples 128x128 = 16384 landmarks\n"); return EXIT_FAILURE; }
    std::vector<float> fixed, moving;
    if (!load_pc8d (o.fixed, fixed) || !load_pc8d (o.moving, moving)) return EXIT_FAILURE;
    try
    {
        if (o.svd) return o.regular ? registerPC<ICPStepConfigT::EIGEN, ICPStepConfigW::REGULAR> (o, fixed, moving)
                                    : registerPC<ICPStepConfigT::EIGEN, ICPStepConfigW::WEIGHTED> (o, fixed, moving);
        return o.regular ? registerPC<ICPStepConfigT::POWER_METHOD, ICPStepConfigW::REGULAR> (o, fixed, moving)
                         : registerPC<ICPStepConfigT::POWER_METHOD, ICPStepConfigW::WEIGHTED> (o, fixed, moving);
    }
    catch (const cl::Error &error)
    {
        std::fprintf (stderr, "%s (%d)\n", error.what (), error.err ());
        return EXIT_FAILURE;
    }
}
