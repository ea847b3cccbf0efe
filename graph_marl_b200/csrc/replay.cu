// replay.cu -- device-resident replay ring: batched insert and index gather.
//
// Reference: src/replaybuffer.py:243-287 (add: 17 per-field copies into ring slot `index`,
// then index = (index+1) % size), :132-187 (_get_transition_batch: fancy-index gather plus
// dtype conversion to float32 / int64 / bool).  Here B transitions (one per env) are
// inserted per call into consecutive slots and the ring never leaves HBM.
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"

namespace gm {

struct ReplayFields {
    gm_replay_field f[GM_REPLAY_MAX_FIELDS];
    int n;
};

// grid.y = field, grid.x strides over (transition, 16 / 8 / 4 / 1-byte unit)
__global__ void replay_insert_kernel(ReplayFields F, int64_t capacity, int64_t index0, const int64_t* __restrict__ index_dev,
                                     int64_t n) {
    const int64_t index = index0 + (index_dev ? *index_dev : 0);
    const gm_replay_field fd = F.f[blockIdx.y];
    const int64_t eb = fd.elem_bytes;
    const int64_t tid0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
    if (fd.convert == 4) {  // int32 source elements -> int8 ring elements (actions)
        const int64_t total = n * eb;
        for (int64_t t = tid0; t < total; t += nthreads) {
            int64_t i = t / eb, e = t - i * eb;
            int64_t slot = (index + i) % capacity;
            ((int8_t*)fd.ring)[slot * eb + e] = (int8_t)((const int32_t*)fd.src)[(fd.broadcast ? 0 : i) * eb + e];
        }
        return;
    }
    const bool two_d = fd.rows > 0;
    const int64_t rb = two_d ? fd.row_bytes : eb;           // contiguous run
    const int64_t runs = two_d ? fd.rows : 1;               // runs per transition
    const uintptr_t src_bits = (uintptr_t)fd.src | (uintptr_t)rb;
    const uintptr_t dst_bits = (uintptr_t)fd.ring | (uintptr_t)eb | (two_d ? ((uintptr_t)fd.ring_pitch | (uintptr_t)fd.ring_offset) : 0);
    const uintptr_t align_bits = src_bits | dst_bits;
    if (align_bits % 16 != 0 && src_bits % 16 == 0 && dst_bits % 8 == 0) {
        // 16-byte aligned source, destination only 8-byte aligned (the graph part of a joint observation row
        // starts 520 bytes into the ring row): one 16-byte load, two 8-byte stores
        const int64_t upr = rb / 16, total = n * runs * upr;
        auto addr2 = [&](int64_t t, const uint4*& s, uint2*& d) {
            int64_t u = t % upr, q = t / upr;
            int64_t row = q % runs, i = q / runs;
            int64_t slot = (index + i) % capacity;
            s = (const uint4*)((const char*)fd.src + ((fd.broadcast ? 0 : i) * runs + row) * rb + u * 16);
            d = (uint2*)((char*)fd.ring + slot * eb + (two_d ? fd.ring_offset + row * fd.ring_pitch : 0) + u * 16);
        };
        int64_t t = tid0;
        for (; t + 3 * nthreads < total; t += 4 * nthreads) {
            const uint4* s[4];
            uint2* d[4];
            uint4 v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) addr2(t + k * nthreads, s[k], d[k]);
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = __ldcs(s[k]);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                __stcs(d[k], make_uint2(v[k].x, v[k].y));
                __stcs(d[k] + 1, make_uint2(v[k].z, v[k].w));
            }
        }
        for (; t < total; t += nthreads) {
            const uint4* s;
            uint2* d;
            addr2(t, s, d);
            const uint4 v = *s;
            d[0] = make_uint2(v.x, v.y);
            d[1] = make_uint2(v.z, v.w);
        }
        return;
    }
    const int unit = (align_bits % 16 == 0) ? 16 : (align_bits % 8 == 0) ? 8 : (align_bits % 4 == 0) ? 4 : 1;
    const int64_t upr = rb / unit;  // units per run
    const int64_t total = n * runs * upr;
    auto addr = [&](int64_t t, const char*& s, char*& d) {
        int64_t u = t % upr, q = t / upr;
        int64_t row = q % runs, i = q / runs;
        int64_t slot = (index + i) % capacity;
        s = (const char*)fd.src + ((fd.broadcast ? 0 : i) * runs + row) * rb + u * unit;
        d = (char*)fd.ring + slot * eb + (two_d ? fd.ring_offset + row * fd.ring_pitch : 0) + u * unit;
    };
    if (unit == 16) {  // bulk of the bytes: four independent 16-byte loads in flight per thread
        int64_t t = tid0;
        for (; t + 3 * nthreads < total; t += 4 * nthreads) {
            const char* s[4];
            char* d[4];
            uint4 v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) addr(t + k * nthreads, s[k], d[k]);
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = __ldcs((const uint4*)s[k]);
#pragma unroll
            for (int k = 0; k < 4; k++) __stcs((uint4*)d[k], v[k]);
        }
        for (; t < total; t += nthreads) {
            const char* s;
            char* d;
            addr(t, s, d);
            *(uint4*)d = *(const uint4*)s;
        }
        return;
    }
    for (int64_t t = tid0; t < total; t += nthreads) {
        const char* s;
        char* d;
        addr(t, s, d);
        if (unit == 8) *(uint2*)d = *(const uint2*)s;
        else if (unit == 4) *(uint32_t*)d = *(const uint32_t*)s;
        else *d = *s;
    }
}

// gather with conversion: one thread per output element
__global__ void replay_sample_kernel(ReplayFields F, const int64_t* __restrict__ indices, int64_t n) {
    const gm_replay_field fd = F.f[blockIdx.y];
    const int in_size = (fd.convert == 0) ? 1 : (fd.convert == 3 ? 2 : 1);
    const int64_t elems = fd.elem_bytes / in_size;  // for convert==0 this is bytes
    const int64_t total = n * elems;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = t / elems, e = t - i * elems;
        int64_t slot = indices[i];
        const char* s = (const char*)fd.ring + slot * fd.elem_bytes;
        switch (fd.convert) {
            case 0: ((char*)fd.dst)[t] = s[e]; break;
            case 1: ((float*)fd.dst)[t] = (float)(((const uint8_t*)s)[e] != 0); break;
            case 2: ((int64_t*)fd.dst)[t] = (int64_t)((const int8_t*)s)[e]; break;
            case 3: ((float*)fd.dst)[t] = __half2float(((const __half*)s)[e]); break;
        }
    }
}

}  // namespace gm

using namespace gm;

extern "C" {

int gm_replay_insert(const gm_replay_field* fields, int32_t n_fields, int64_t capacity, int64_t index, const int64_t* index_dev,
                     int64_t n, void* stream) {
    GM_CHECK_ARG(fields && n_fields > 0 && n_fields <= GM_REPLAY_MAX_FIELDS, "bad field count %d", n_fields);
    GM_CHECK_ARG(capacity > 0 && n >= 0 && n <= capacity && index >= 0 && (index_dev != nullptr || index < capacity), "bad ring arguments");
    if (n == 0) return GM_OK;
    ReplayFields F;
    F.n = n_fields;
    int64_t maxb = 0;
    for (int i = 0; i < n_fields; i++) {
        GM_CHECK_ARG(fields[i].ring && fields[i].src && fields[i].elem_bytes > 0, "field %d: null ring/src", i);
        GM_CHECK_ARG(fields[i].convert == 0 || fields[i].convert == 4, "field %d: insert convert %d", i, fields[i].convert);
        GM_CHECK_ARG(fields[i].rows == 0 ||
                         (fields[i].row_bytes > 0 && fields[i].ring_offset >= 0 &&
                          fields[i].ring_offset + (fields[i].rows - 1) * fields[i].ring_pitch + fields[i].row_bytes <= fields[i].elem_bytes),
                     "field %d: 2-D block does not fit the ring element", i);
        F.f[i] = fields[i];
        maxb = max(maxb, fields[i].elem_bytes * n);
    }
    int64_t blocks = std::min<int64_t>((maxb / 64 + 255) / 256 + 1, 148 * 4);
    dim3 grid((unsigned)blocks, n_fields);
    {
        ProfileScope prof(PROF_REPLAY, (cudaStream_t)stream);
        replay_insert_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(F, capacity, index, index_dev, n);
    }
    GM_LAUNCH_CHECK();
    return GM_OK;
}

int gm_replay_sample(const gm_replay_field* fields, int32_t n_fields, const int64_t* indices, int64_t n, void* stream) {
    GM_CHECK_ARG(fields && n_fields > 0 && n_fields <= GM_REPLAY_MAX_FIELDS && indices, "bad sample arguments");
    if (n == 0) return GM_OK;
    ReplayFields F;
    F.n = n_fields;
    int64_t maxe = 0;
    for (int i = 0; i < n_fields; i++) {
        GM_CHECK_ARG(fields[i].ring && fields[i].dst && fields[i].elem_bytes > 0, "field %d: null ring/dst", i);
        GM_CHECK_ARG(fields[i].convert >= 0 && fields[i].convert <= 3, "field %d: convert %d", i, fields[i].convert);
        F.f[i] = fields[i];
        maxe = max(maxe, fields[i].elem_bytes * n);
    }
    int64_t blocks = std::min<int64_t>((maxe + 255) / 256, 148 * 8);
    dim3 grid((unsigned)blocks, n_fields);
    replay_sample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(F, indices, n);
    GM_LAUNCH_CHECK();
    return GM_OK;
}

}  // extern "C"
