// fd_decode.cu — RetinaFace head decode for all strides and the whole batch in one launch, plus the result
// finalisation (keep-gather + rescale): the three-kernel path (decode_kernel -> nms_cta_kernel -> finalize_kernel) that
// the fused kernel (fd_detect_fused.cu) replaces in the batched pipeline; kept for geometries without 128-bit score rows,
// for images the fused kernel defers, and as the A/B reference (FD_NO_FUSED=1).  Arithmetic in fd_decode.cuh.
//
// Replaces face_detection.rs:319-408 (per-stride decode), rcnn/anchors.rs:3-21 (anchor plane, recomputed on the fly
// from 6 base anchors), face_detection.rs:516-570 (bbox_pred / landmark_pred), bbox_transform.rs:27-45 (clip_boxes),
// the >= threshold compaction (:374-379) and, in finalize, the keep-gather (:432-464) and _postprocess (:473-493).
//
// Layout: heads are the network's NCHW tensors with a batch dimension, read in place (no NHWC re-layout): a warp
// walks consecutive (h,w) positions of one channel plane, so every load is a coalesced 128-byte line.  The score
// planes are read for every anchor; the 14 regression planes only where score >= thr.  Candidates are written
// sparsely, indexed by their global anchor id (order 32|16|8, then (h,w,a) — face_detection.rs:410), so the later
// sort key (score desc, anchor id asc) reproduces the reference's stable ordering without an ordered compaction.
#include "fd_internal.cuh"
#include "fd_decode.cuh"

namespace fd {

// decodes anchor (s, local, a) of image b and writes its candidate record
__device__ __forceinline__ void decode_one(const DecodeCfg &c, const HeadPtrs &hp, int b, int s, int local, int a, float score,
                                           int slot, u64 *__restrict__ keys, float4 *__restrict__ cand_box,
                                           float *__restrict__ cand_rec) {
    const size_t img_base = (size_t)b * c.total_anchors;
    const int id = c.anchor_off[s] + local * c.A + a;
    const AnchorGeo g = anchor_geo(c, s, local, a);
    cand_box[img_base + id] = decode_box(c, hp, b, s, local, a, g);
    float rec[CAND_REC];
    decode_landmarks(c, hp, b, s, local, a, g, rec);
    rec[10] = score;
    rec[11] = 0.0f;
    float4 *dst = reinterpret_cast<float4 *>(cand_rec + (img_base + id) * CAND_REC);
    dst[0] = make_float4(rec[0], rec[1], rec[2], rec[3]);
    dst[1] = make_float4(rec[4], rec[5], rec[6], rec[7]);
    dst[2] = make_float4(rec[8], rec[9], rec[10], rec[11]);
    keys[img_base + slot] = ((u64)desc_key(score) << 32) | (unsigned)id;
}

// VEC consecutive positions of one stride per thread (VEC == 4: 128-bit loads of the score planes; needs every H*W to be
// a multiple of 4 and 16-byte aligned tensors, else VEC == 1).  One atomicAdd per warp reserves the key slots.
template <int VEC>
__global__ void __launch_bounds__(256) decode_kernel(DecodeCfg c, HeadPtrs hp, float conf_thr, u64 *__restrict__ keys,
                                                     float4 *__restrict__ cand_box, float *__restrict__ cand_rec,
                                                     int *__restrict__ counts, int *__restrict__ status) {
    const int b = blockIdx.y;
    const int pos = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    const int lane = threadIdx.x & 31;
    const bool active = pos < c.total_pos;
    const int s = active ? stride_of_pos(c, pos) : 0;
    const int local = active ? pos - c.pos_off[s] : 0;
    const int hw = c.fh[s] * c.fw[s];
    const int A = c.A;
    const float *sc = hp.p[3 * s] + (size_t)b * 2 * A * hw;
    float score[FD_MAX_ANCHORS][VEC];
    unsigned pass = 0;  // bit a*VEC+v
    if (active) {
#pragma unroll
        for (int a = 0; a < FD_MAX_ANCHORS; ++a) {
            if (a >= A) break;
            const float *p = sc + (size_t)(A + a) * hw + local;  // fg scores are channels A.. (face_detection.rs:322)
            if (VEC == 4) {
                const float4 q = __ldg(reinterpret_cast<const float4 *>(p));
                score[a][0] = q.x; score[a][1 % VEC] = q.y; score[a][2 % VEC] = q.z; score[a][3 % VEC] = q.w;
            } else {
                score[a][0] = __ldg(p);
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float sv = score[a][v];
                if (sv >= conf_thr) pass |= 1u << (a * VEC + v);         // face_detection.rs:375 (a NaN score fails `>=`: dropped)
            }
        }
    }
    const int cnt = __popc(pass);
    if (__ballot_sync(0xffffffffu, cnt != 0) == 0) return;  // common case: nothing above the threshold in this warp
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int nb = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += nb;
    }
    int base = 0;
    if (lane == 31) base = atomicAdd(&counts[b], incl);
    base = __shfl_sync(0xffffffffu, base, 31);
    int slot = base + incl - cnt;
    while (pass) {
        const int bit = __ffs(pass) - 1;
        pass &= pass - 1;
        const int a = bit / VEC, v = bit - a * VEC;
        float sv = 0.0f;
#pragma unroll
        for (int aa = 0; aa < FD_MAX_ANCHORS; ++aa)
#pragma unroll
            for (int vv = 0; vv < VEC; ++vv)
                if (aa == a && vv == v) sv = score[aa][vv];
        decode_one(c, hp, b, s, local + v, a, sv, slot++, keys, cand_box, cand_rec);
    }
}

// One CTA per image: exclusive offset over the kept counts, then gather + rescale (division, face_detection.rs:477-483).
__global__ void __launch_bounds__(256) finalize_kernel(int B, int TA, const int *__restrict__ keep_count,
                                                       const int *__restrict__ keep_src, const float4 *__restrict__ cand_box,
                                                       const float *__restrict__ cand_rec, const float *__restrict__ det_scale,
                                                       int *__restrict__ offsets, float *__restrict__ out_det,
                                                       float *__restrict__ out_lmk, int *__restrict__ out_frame_idx,
                                                       int *__restrict__ status) {
    __shared__ int red[8];
    const int b = blockIdx.x, tid = threadIdx.x;
    int part = 0;
    for (int i = tid; i < b; i += 256) part += max(keep_count[i], 0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    int offset = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) offset += red[k];
    const int M = max(keep_count[b], 0);
    if (tid == 0) {
        offsets[b] = offset;
        if (b == B - 1) {
            offsets[B] = offset + M;
            status[2] = offset + M;
        }
    }
    const float ds = det_scale[b];
    const size_t img_base = (size_t)b * TA;
    for (int m = tid; m < M; m += 256) {
        const int id = keep_src[img_base + m];
        const float4 bx = cand_box[img_base + id];
        const float *rec = cand_rec + (img_base + id) * CAND_REC;
        float *d = out_det + (size_t)(offset + m) * 5;
        d[0] = __fdiv_rn(bx.x, ds);
        d[1] = __fdiv_rn(bx.y, ds);
        d[2] = __fdiv_rn(bx.z, ds);
        d[3] = __fdiv_rn(bx.w, ds);
        d[4] = rec[10];
        float *l = out_lmk + (size_t)(offset + m) * 10;
#pragma unroll
        for (int k = 0; k < 10; ++k) l[k] = __fdiv_rn(rec[k], ds);
        out_frame_idx[offset + m] = b;
    }
}

int decode_launch(fd_ctx *ctx, const float *const *heads_dev, int B, float conf_thr) {
    HeadPtrs hp;
    for (int i = 0; i < 3 * FD_MAX_STRIDES; ++i) hp.p[i] = i < 3 * ctx->dcfg.n_strides ? heads_dev[i] : nullptr;
    bool vec4 = true;  // 128-bit score loads need H*W % 4 == 0 for every stride and 16-byte aligned score tensors
    for (int st = 0; st < ctx->dcfg.n_strides; ++st)
        vec4 = vec4 && ((ctx->dcfg.fh[st] * ctx->dcfg.fw[st]) % 4 == 0) && (reinterpret_cast<uintptr_t>(heads_dev[3 * st]) % 16 == 0);
    if (vec4) {
        dim3 grid((ctx->dcfg.total_pos / 4 + 255) / 256, B);
        decode_kernel<4><<<grid, 256, 0, ctx->stream>>>(ctx->dcfg, hp, conf_thr, ctx->cand_keys.as<u64>(), ctx->cand_box.as<float4>(),
                                                       ctx->cand_lmk.as<float>(), ctx->cand_count.as<int>(), ctx->status());
    } else {
        dim3 grid((ctx->dcfg.total_pos + 255) / 256, B);
        decode_kernel<1><<<grid, 256, 0, ctx->stream>>>(ctx->dcfg, hp, conf_thr, ctx->cand_keys.as<u64>(), ctx->cand_box.as<float4>(),
                                                       ctx->cand_lmk.as<float>(), ctx->cand_count.as<int>(), ctx->status());
    }
    FD_LAUNCH_CHECK_NAMED(ctx, "decode_kernel");
    return FD_OK;
}

int finalize_launch(fd_ctx *ctx, int B) {
    finalize_kernel<<<B, 256, 0, ctx->stream>>>(B, ctx->dcfg.total_anchors, ctx->keep_count.as<int>(), ctx->keep_src.as<int>(),
                                                ctx->cand_box.as<float4>(), ctx->cand_lmk.as<float>(),
                                                ctx->det_scale_dev.as<float>(), ctx->out_offsets.as<int>(),
                                                ctx->out_det.as<float>(), ctx->out_lmk.as<float>(),
                                                ctx->out_frame_idx.as<int>(), ctx->status());
    FD_LAUNCH_CHECK_NAMED(ctx, "finalize_kernel");
    return FD_OK;
}

}  // namespace fd
