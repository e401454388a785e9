// icp_multi.cu -- single-process multi-GPU registration of independent frame pairs ("one host thread + one
// context / stream set per GPU", contiguous pair blocks, no collective on the hot path, only the 8-float poses come back).
//
// A registration does not shard (3-4 global reductions per iteration over ~3 us of work), so the unit of distribution is
// the frame pair: device d owns the pairs [first_d, first_d + count_d) end to end -- upload, buildRBC, n iterations,
// pose read-back -- through its own icp_ctx + icp_batch (the engine of icp_batch.cu, unchanged).  The host side is one
// std::thread per device for the duration of a call; the devices never talk to each other.
#include "icp_common.cuh"
#include <string.h>
#include <thread>

struct icp_multi
{
    uint32_t n_pairs = 0, m = 0;
    std::vector<int> devices;
    std::vector<icp_ctx *> ctx;
    std::vector<icp_batch *> batch;
    std::vector<uint32_t> first, count;      // contiguous block of every device (remainders go to the lowest devices)
};

extern "C" void icp_multi_destroy(icp_multi *mg)
{
    if (!mg) return;
    for (icp_batch *b : mg->batch) icp_batch_destroy(b);
    for (icp_ctx *c : mg->ctx) icp_ctx_destroy(c);
    delete mg;
}

extern "C" int icp_multi_create(int n_devices, const int *devices, int rot_cfg, int w_cfg, uint32_t n_pairs, uint32_t m, uint32_t nr,
                                float alpha, float c, uint32_t lm_w, uint32_t lm_h, icp_multi **out)
{
    if (!out || n_pairs == 0 || n_devices < 0) { icp_set_error("icp_multi_create: bad argument"); return ICP_ERR_ARG; }
    int visible = 0;
    cudaError_t e = cudaGetDeviceCount(&visible);
    if (e != cudaSuccess || visible == 0)
    {
        icp_set_error("icp_multi_create: no CUDA device available (%s); libicp_b200 has no CPU fallback",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return ICP_ERR_CUDA;
    }
    if (n_devices == 0) n_devices = visible;                                  // all visible GPUs
    if ((uint32_t)n_devices > n_pairs) n_devices = (int)n_pairs;              // at least one pair per device
    icp_multi *mg = new icp_multi();
    mg->n_pairs = n_pairs; mg->m = m;
    const uint32_t base = n_pairs / (uint32_t)n_devices, rem = n_pairs % (uint32_t)n_devices;
    uint32_t first = 0;
    for (int d = 0; d < n_devices; ++d)
    {
        const int dev = devices ? devices[d] : d;
        if (dev < 0 || dev >= visible)
        {
            icp_set_error("icp_multi_create: device %d out of range [0,%d)", dev, visible);
            icp_multi_destroy(mg);
            return ICP_ERR_ARG;
        }
        const uint32_t cnt = base + ((uint32_t)d < rem ? 1u : 0u);
        icp_ctx *cx = nullptr;
        int rc = icp_ctx_create(dev, nullptr, &cx);
        if (rc == ICP_OK) mg->ctx.push_back(cx);
        icp_batch *b = nullptr;
        if (rc == ICP_OK) rc = icp_batch_create(cx, rot_cfg, w_cfg, cnt, m, nr, alpha, c, lm_w, lm_h, &b);
        if (rc != ICP_OK) { icp_multi_destroy(mg); return rc; }
        mg->batch.push_back(b);
        mg->devices.push_back(dev);
        mg->first.push_back(first);
        mg->count.push_back(cnt);
        first += cnt;
    }
    *out = mg;
    return ICP_OK;
}

extern "C" int icp_multi_devices(icp_multi *mg) { return mg ? (int)mg->devices.size() : 0; }

extern "C" int icp_multi_pair_range(icp_multi *mg, int index, int *device, uint32_t *first, uint32_t *count)
{
    if (!mg || index < 0 || index >= (int)mg->devices.size()) { icp_set_error("icp_multi_pair_range: bad index"); return ICP_ERR_ARG; }
    if (device) *device = mg->devices[index];
    if (first) *first = mg->first[index];
    if (count) *count = mg->count[index];
    return ICP_OK;
}

extern "C" int icp_multi_register_host(icp_multi *mg, const float *h_F, const float *h_M, uint32_t n_iters, float *h_T8)
{
    if (!mg || !h_F || !h_M || !h_T8 || n_iters == 0) { icp_set_error("icp_multi_register_host: bad argument"); return ICP_ERR_ARG; }
    const size_t nd = mg->devices.size();
    const size_t per = (size_t)mg->m * 8;
    std::vector<int> rc(nd, ICP_OK);
    std::vector<std::string> msg(nd);
    auto work = [&](size_t d)
    {
        // blocking host-buffer entry of the device's own batch: sliced h2d overlapped with the registration of the previous slice
        rc[d] = icp_batch_register_host(mg->batch[d], h_F + mg->first[d] * per, h_M + mg->first[d] * per, n_iters, 0, h_T8 + (size_t)mg->first[d] * 8);
        if (rc[d] != ICP_OK) msg[d] = icp_last_error();       // the error string is thread local: carry it to the caller
    };
    std::vector<std::thread> th;
    for (size_t d = 1; d < nd; ++d) th.emplace_back(work, d);
    work(0);
    for (std::thread &t : th) t.join();
    for (size_t d = 0; d < nd; ++d)
        if (rc[d] != ICP_OK) { icp_set_error("device %d: %s", mg->devices[d], msg[d].c_str()); return rc[d]; }
    return ICP_OK;
}

// one-shot convenience: create on n_devices GPUs (0 = all visible), register, destroy
extern "C" int icp_multi_register_host_once(int n_devices, int rot_cfg, int w_cfg, uint32_t n_pairs, uint32_t m, uint32_t nr, float alpha, float c,
                                            const float *h_F, const float *h_M, uint32_t n_iters, float *h_T8)
{
    icp_multi *mg = nullptr;
    ICP_CHECK(icp_multi_create(n_devices, nullptr, rot_cfg, w_cfg, n_pairs, m, nr, alpha, c, 0, 0, &mg));
    const int rc = icp_multi_register_host(mg, h_F, h_M, n_iters, h_T8);
    icp_multi_destroy(mg);
    return rc;
}
