// green_ctx.cu - SM partitions for the two streams of the chunk pipeline (chb_api.cu, convolutions_all).
//
// On several GPUs the kernels whose stores are the pencil transposes (zfwd, xpass) wait for NVLink, the others for
// HBM or the FP64 pipe.  Two plain streams do not make them overlap: each kernel's grid fills every SM (and its whole
// shared memory), so the block scheduler runs them one after the other.  CUDA green contexts (driver API, CUDA >= 12.4)
// give each stream a fixed, disjoint set of SMs: partition A runs the x-pass, partition B the z-passes and the RHS
// assembly of the neighbouring chunks, side by side for the whole sweep (chb_api.cu, convolutions_all).  This is the role of the reference's nonblockingXZ variant
// (mpi_transpose.f90:149-168: MPI_IAlltoall progressing under the next plane's FFTs).
//
// The device is split into the smallest groups the driver offers (2 SMs when SM co-scheduling is ignored - no kernel
// here uses clusters -, else 8); the first groups form A, the others and the remainder of the split B.  libcuda is reached through dlopen, like NCCL, so that the library links against the runtime only.
// Everything here is optional: any failure leaves the handle on plain streams (chb_green_create returns non-zero).
#include <cuda.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "chb_internal.h"

#if defined(CHB_HOST_EMUL)
int chb_green_create(chb_handle_s*, int) { return 1; }
void chb_green_destroy(chb_handle_s*) {}
#else
namespace {
struct DriverApi {
    void* lib = nullptr;
    CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
    CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int,
                                          unsigned int) = nullptr;
    CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
    CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
    CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
    bool ok = false;
};
DriverApi g_drv;

bool load_driver() {
    if (g_drv.lib) return g_drv.ok;
    g_drv.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!g_drv.lib) return false;
#define LOAD(field, sym) *(void**)(&g_drv.field) = dlsym(g_drv.lib, sym)
    LOAD(DeviceGet, "cuDeviceGet");
    LOAD(DeviceGetDevResource, "cuDeviceGetDevResource");
    LOAD(DevSmResourceSplitByCount, "cuDevSmResourceSplitByCount");
    LOAD(DevResourceGenerateDesc, "cuDevResourceGenerateDesc");
    LOAD(GreenCtxCreate, "cuGreenCtxCreate");
    LOAD(GreenCtxDestroy, "cuGreenCtxDestroy");
    LOAD(GreenCtxStreamCreate, "cuGreenCtxStreamCreate");
#undef LOAD
    g_drv.ok = g_drv.DeviceGet && g_drv.DeviceGetDevResource && g_drv.DevSmResourceSplitByCount &&
               g_drv.DevResourceGenerateDesc && g_drv.GreenCtxCreate && g_drv.GreenCtxDestroy && g_drv.GreenCtxStreamCreate;
    return g_drv.ok;
}
}  // namespace

// sms_a > 0: SMs of partition A (rounded to whole groups); sms_a < 0: the default share, 54 % of the device.
int chb_green_create(chb_handle_s* h, int sms_a) {
    if (!load_driver()) return 1;
    cudaFree(0);   // the primary context exists and is current
    CUdevice dev;
    if (g_drv.DeviceGet(&dev, h->device) != CUDA_SUCCESS) return 1;
    CUdevResource all;
    if (g_drv.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return 1;
    const int total = (int)all.sm.smCount;
    // groups of 8 SMs aligned to the GPC structure leave 28 of the 148 SMs of a B200 in the remainder; none of our
    // kernels uses thread-block clusters, so the split may ignore SM co-scheduling and use the finest granularity
    unsigned int ng = 0;
    unsigned int flags = CU_DEV_SM_RESOURCE_SPLIT_IGNORE_SM_COSCHEDULING, gran = 2;
    if (g_drv.DevSmResourceSplitByCount(nullptr, &ng, &all, nullptr, flags, gran) != CUDA_SUCCESS || ng < 2) {
        flags = 0; gran = 8; ng = 0;
        if (g_drv.DevSmResourceSplitByCount(nullptr, &ng, &all, nullptr, flags, gran) != CUDA_SUCCESS || ng < 2) return 1;
    }
    std::vector<CUdevResource> grp(ng + 1);
    CUdevResource rem;
    memset(&rem, 0, sizeof(rem));
    if (g_drv.DevSmResourceSplitByCount(grp.data(), &ng, &all, &rem, flags, gran) != CUDA_SUCCESS || ng < 2) return 1;
    const int per = (int)grp[0].sm.smCount;
    if (sms_a < 0) sms_a = (int)(0.54 * total + 0.5);
    int ka = (sms_a + per / 2) / per;
    if (ka < 1) ka = 1;
    if (ka > (int)ng - 1) ka = (int)ng - 1;
    // partition B: the other groups and what the split left over
    unsigned int nb = ng - (unsigned)ka;
    int sms_b = (int)nb * per;
    if (rem.type == CU_DEV_RESOURCE_TYPE_SM && rem.sm.smCount > 0) {
        grp[ng] = rem;
        nb += 1;
        sms_b += (int)rem.sm.smCount;
    }
    CUdevResourceDesc da, db;
    if (g_drv.DevResourceGenerateDesc(&da, grp.data(), (unsigned)ka) != CUDA_SUCCESS) return 1;
    if (g_drv.DevResourceGenerateDesc(&db, grp.data() + ka, nb) != CUDA_SUCCESS) {
        // without the remainder
        nb = ng - (unsigned)ka;
        sms_b = (int)nb * per;
        if (g_drv.DevResourceGenerateDesc(&db, grp.data() + ka, nb) != CUDA_SUCCESS) return 1;
    }
    CUgreenCtx ga = nullptr, gb = nullptr;
    if (g_drv.GreenCtxCreate(&ga, da, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return 1;
    if (g_drv.GreenCtxCreate(&gb, db, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) {
        g_drv.GreenCtxDestroy(ga);
        return 1;
    }
    CUstream sa = nullptr, sb = nullptr;
    if (g_drv.GreenCtxStreamCreate(&sa, ga, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS ||
        g_drv.GreenCtxStreamCreate(&sb, gb, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) {
        if (sa) cudaStreamDestroy((cudaStream_t)sa);
        g_drv.GreenCtxDestroy(ga);
        g_drv.GreenCtxDestroy(gb);
        return 1;
    }
    h->green[0] = ga;
    h->green[1] = gb;
    h->green_sms[0] = ka * per;
    h->green_sms[1] = sms_b;
    h->sA = (cudaStream_t)sa;
    h->sB = (cudaStream_t)sb;
    if (getenv("CHB_VERBOSE"))
        fprintf(stderr, "[channel_b200] rank %d: green contexts, %d SMs for the transposes + %d SMs for the local kernels (of %d)\n",
                h->g.rank, h->green_sms[0], h->green_sms[1], total);
    return 0;
}

// after the streams have been destroyed
void chb_green_destroy(chb_handle_s* h) {
    for (int i = 0; i < 2; ++i)
        if (h->green[i]) {
            g_drv.GreenCtxDestroy((CUgreenCtx)h->green[i]);
            h->green[i] = nullptr;
        }
}
#endif
