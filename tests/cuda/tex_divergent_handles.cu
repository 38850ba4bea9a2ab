// tex_divergent_handles.cu — does a TEX with a per-lane (divergent) bindless handle return the right image on sm_100a?
//
// shade.cuh's sample_texture walks the list of real images with a warp-uniform counter because the first version of the
// wavefront shading kernel — `tex2D<uchar4>(S.textures[index].obj, ..)` with a per-lane `index` — returned texels of the
// wrong image for some lanes on B200 (nvcc 12.9, -arch=sm_100a) when the lanes of a warp held different handles inside
// divergent code.  This program is the stand-alone check of that pattern: N single-colour images (image k holds the
// value k in every texel), every lane picks its image by a hash, fetches inside data-dependent control flow, and the
// host counts lanes whose texel is not their image's value.  Four variants:
//   0  direct:     tex2D(handle_of_lane)                      (the compiler serialises over distinct handles)
//   1  elected:    loop { leader = ffs(todo); k = shfl(index, leader); if (index == k) tex2D(handle[k]) }
//   2  counter:    for j in 0..N (uniform counter): if (index == j) tex2D(handle[j])        (what shade.cuh ships)
//   3  present:    as 2, but j only runs over the indices present in the warp (match-any mask reduced with redux.or)
// usage: tex_divergent_handles [images=96]     exit code 0; prints mismatches per variant.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 2; } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int VARIANT>
__global__ void k_fetch(const cudaTextureObject_t* __restrict__ handles, uint32_t n_images, uint32_t* __restrict__ got, uint32_t* __restrict__ want,
                        uint32_t rounds) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc_got = 0, acc_want = 0;
    for (uint32_t r = 0; r < rounds; r++) {
        const uint32_t h = hash32(tid * 977u + r * 131071u);
        // data-dependent control flow around the fetch, like a traversal loop calling the any-hit
        const uint32_t trips = 1u + (h >> 28);
        for (uint32_t t = 0; t < trips; t++) {
            const uint32_t hh = hash32(h + t);
            if ((hh & 3u) == 0u) continue;  // this lane has no candidate in this trip
            const uint32_t index = hh % n_images;
            uint32_t v = 0xFFFFFFFFu;
            const float x = (float)((hh >> 8) & 3u) + 0.5f, y = (float)((hh >> 10) & 3u) + 0.5f;
            if (VARIANT == 0) {
                v = tex2D<uchar4>(handles[index], x, y).x;
            } else if (VARIANT == 1) {
                const uint32_t active = __activemask();
                uint32_t todo = active;
                while (todo) {
                    const int leader = __ffs(todo) - 1;
                    const uint32_t k = __shfl_sync(active, index, leader);
                    const bool mine = index == k;
                    if (mine) v = tex2D<uchar4>(handles[k], x, y).x;
                    todo &= ~__ballot_sync(active, mine);
                }
            } else if (VARIANT == 2) {
                for (uint32_t j = 0; j < n_images; j++)
                    if (j == index) v = tex2D<uchar4>(handles[j], x, y).x;
            } else {
                const uint32_t active = __activemask();
                for (uint32_t word = 0; word * 32u < n_images; word++) {
                    uint32_t present = __reduce_or_sync(active, (index >> 5) == word ? 1u << (index & 31u) : 0u);
                    for (uint32_t j = word * 32u; present; j++, present >>= 1)
                        if ((present & 1u) && j == index) v = tex2D<uchar4>(handles[j], x, y).x;
                }
            }
            acc_got = acc_got * 31u + v;
            acc_want = acc_want * 31u + index;
        }
    }
    got[tid] = acc_got;
    want[tid] = acc_want;
}

int main(int argc, char** argv) {
    const uint32_t n_images = argc > 1 ? (uint32_t)atoi(argv[1]) : 96u;
    std::vector<cudaArray_t> arrays(n_images);
    std::vector<cudaTextureObject_t> objs(n_images);
    for (uint32_t k = 0; k < n_images; k++) {
        cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
        CK(cudaMallocArray(&arrays[k], &cd, 4, 4));
        std::vector<uchar4> texels(16, make_uchar4((unsigned char)k, 0, 0, 255));
        CK(cudaMemcpy2DToArray(arrays[k], 0, 0, texels.data(), 16, 16, 4, cudaMemcpyHostToDevice));
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = arrays[k];
        cudaTextureDesc td = {};
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeElementType;
        CK(cudaCreateTextureObject(&objs[k], &rd, &td, nullptr));
    }
    cudaTextureObject_t* d_handles;
    CK(cudaMalloc(&d_handles, sizeof(cudaTextureObject_t) * n_images));
    CK(cudaMemcpy(d_handles, objs.data(), sizeof(cudaTextureObject_t) * n_images, cudaMemcpyHostToDevice));
    const uint32_t threads = 148u * 8u * 128u, rounds = 64u;
    uint32_t *d_got, *d_want;
    CK(cudaMalloc(&d_got, threads * 4));
    CK(cudaMalloc(&d_want, threads * 4));
    std::vector<uint32_t> got(threads), want(threads);
    int rc = 0;
    for (int variant = 0; variant < 4; variant++) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        switch (variant) {
            case 0: k_fetch<0><<<threads / 128, 128>>>(d_handles, n_images, d_got, d_want, rounds); break;
            case 1: k_fetch<1><<<threads / 128, 128>>>(d_handles, n_images, d_got, d_want, rounds); break;
            case 2: k_fetch<2><<<threads / 128, 128>>>(d_handles, n_images, d_got, d_want, rounds); break;
            default: k_fetch<3><<<threads / 128, 128>>>(d_handles, n_images, d_got, d_want, rounds); break;
        }
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        CK(cudaMemcpy(got.data(), d_got, threads * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(want.data(), d_want, threads * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (uint32_t i = 0; i < threads; i++) bad += got[i] != want[i];
        const char* names[] = {"direct per-lane handle", "elected leader + shuffle", "uniform counter over all images", "uniform counter over present images"};
        printf("variant %d (%s): %zu of %u threads saw a wrong texel, %.3f ms\n", variant, names[variant], bad, threads, ms);
    }
    return rc;
}
