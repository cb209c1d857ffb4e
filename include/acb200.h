/*
 * acb200.h -- thin C-ABI CUDA layer of the B200 (sm_100a) backend for Anime4KCPP v3's
 * CNN upscaling hot path.
 *
 * This is the seam the C++ host (`ac::core::Processor` re-creation, src/host/) talks to, and the
 * boundary a reference maintainer would bind instead of core/src/processor/cuda/{Kernel.cu,
 * CUDAProcessor.cpp}.  Plain pointers and sizes only; no C++/torch types.  Entry points and what
 * they replace in the reference:
 *
 *   acb200_device_count / acb200_device_info   <- CUDAProcessor.cpp:68-83 (ContextList), :882-892 (info<CUDA>)
 *   acb200_model_create / _destroy             <- CUDAProcessor.cpp:291-326 (per-layer cudaMalloc+cudaMemcpy of
 *                                                 model.kernel(l)/bias(l)/alpha(l))
 *   acb200_session_create / _destroy           <- CUDAProcessor.cpp:96-228 (per-thread stream, pool allocator,
 *                                                 scratch images), :374-377
 *   acb200_process_host                        <- Processor.cpp:199-276 (colour split, 2x passes, chroma
 *                                                 resize, merge) + CUDAProcessor.cpp:383-422/:451-490/:520-566
 *                                                 (H2D, layer launches, D2H, sync) in ONE submission
 *   acb200_process_device                      <- same, for frames already resident in HBM (no copies)
 *   acb200_session_sync / acb200_error_string  <- CUDAProcessor.cpp:260-267 (sticky cudaError_t + string)
 *
 * Image memory layout is the reference's `ac::core::Image` (core/include/AC/Core/Image.hpp:359-419):
 * interleaved HWC rows, `stride` in BYTES, element type code = (kind << 8) | bytes.
 *
 * Every function returns 0 on success or a negative code (ACB200_E*); nothing throws or aborts.
 * There is no CPU fallback: without a usable CUDA device every compute entry fails with
 * ACB200_ENODEVICE.
 */
#ifndef ACB200_H
#define ACB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#   define ACB200_API __attribute__((visibility("default")))
#else
#   define ACB200_API
#endif

/* element types, core/include/AC/Core/Image.hpp:365-370 */
#define ACB200_UINT8   0x001
#define ACB200_UINT16  0x002
#define ACB200_FLOAT16 0x202
#define ACB200_FLOAT32 0x204

/* model families on the hot path, core/src/processor/Processor.cpp:26-187 */
#define ACB200_FAMILY_ACNET_LEGACY 0   /* model::ACNetLegacy: ReLU, 2x2 deconvolution tail */
#define ACB200_FAMILY_ACNET        1   /* model::ACNet<8>:   PReLU, pixel-shuffle + nearest residual tail */
#define ACB200_FAMILY_ARNET        2   /* model::ARNet<8>:   residual blocks, 1x1 fuse, pixel-shuffle tail */
#define ACB200_FAMILY_ARTCNN       3   /* model::ArtCNN<16/32>: ReLU convs, long skip, pixel-shuffle tail (no luma residual) */
#define ACB200_FAMILY_FSRCNNX      4   /* model::FSRCNNX<8/16>: 5x5 head, PReLU convs, 1x1 + skip, pixel-shuffle tail */

#define ACB200_OK          0
#define ACB200_EINVAL     (-22)   /* bad argument / unsupported shape */
#define ACB200_ENODEVICE  (-19)   /* no CUDA device, or device index out of range */
#define ACB200_ENOMEM     (-12)
#define ACB200_ECUDA      (-256)  /* a CUDA call failed; acb200_session_error()/acb200_last_error() has the text */

typedef struct acb200_model acb200_model;
typedef struct acb200_session acb200_session;

ACB200_API int acb200_device_count(void);
/* name: device name (NUL-terminated, truncated to name_len); vram in bytes; cc = major*10+minor */
ACB200_API int acb200_device_info(int device, char* name, int name_len, size_t* vram_bytes, int* cc, int* sm_count, int* clock_khz);

/*
 * Flat fp32 arrays exactly as the reference's model objects expose them
 * (core/include/AC/Core/Model/ACNet.hpp:34-60,78-115, ARNet.hpp:32-71): kernels `[cout][ky*3+kx][cin]`
 * per layer, layers concatenated.  Lengths are validated against family/blocks.
 * The model is host-side and device-independent; sessions upload it lazily per device.
 */
ACB200_API int acb200_model_create(int family, int blocks,
                                   const float* kernels, int n_kernels,
                                   const float* biases, int n_biases,
                                   const float* alphas, int n_alphas,
                                   acb200_model** out);
/*
 * The reference's other model families (core/include/AC/Core/Model/{ArtCNN,FSRCNNX}.hpp; replaces Processor::create<CUDA,
 * model::ArtCNN<F>> / <CUDA, model::FSRCNNX<F>>, core/src/processor/cuda/CUDAProcessor.cpp:574-880): `features` = F
 * (ArtCNN 16 / 32, FSRCNNX 8 / 16), arrays exactly as model.kernel() / bias() / alpha() hand them out.  These families run
 * per-layer fp32 kernels in the reference FMA-backend order (bit-identical); the engine selection does not apply.
 * With family 0..2 and features == 8 this is acb200_model_create.
 */
ACB200_API int acb200_model_create_wide(int family, int features, int blocks,
                                        const float* kernels, int n_kernels, const float* biases, int n_biases,
                                        const float* alphas, int n_alphas, acb200_model** out);
ACB200_API void acb200_model_destroy(acb200_model* model);

/*
 * A session owns one CUDA stream, its scratch planes in HBM and pinned staging buffers on `device`.
 * One session per calling thread (the reference keeps the same state per thread,
 * CUDAProcessor.cpp:374-377); a session must not be used from two threads at once.
 */
ACB200_API int acb200_session_create(int device, acb200_session** out);
ACB200_API void acb200_session_destroy(acb200_session* session);
ACB200_API int acb200_session_device(const acb200_session* session);
/* sticky error text of the last failure on this session ("NO ERROR" when none) */
ACB200_API const char* acb200_session_error(const acb200_session* session);
ACB200_API void acb200_session_clear_error(acb200_session* session);

/*
 * Processor::process(src, dst, factor) for HOST images: H2D, RGB->YUV split (c = 3/4), ceil(log2(factor))
 * 2x luma passes, Catmull-Rom chroma resize by `factor`, YUV->RGB merge, D2H, stream sync.
 * dst must be preallocated: (int)(w*factor) x (int)(h*factor) x c, same element type.
 * Any `factor` >= 1 (<= 64): for factors that are not powers of two the luma is down-scaled after the passes by
 * fxy = factor / 2^power with the Catmull-Rom filter (Processor.cpp:203-204, 237, 249); factor < 1 -> ACB200_EINVAL.
 * Video frames (acb200_process_frame_*) take the same factors; row bands (acb200_process_host_band) powers of two only.
 */
ACB200_API int acb200_process_host(acb200_session* session, const acb200_model* model,
                                   const void* src, int w, int h, int c, int src_stride, int elem_type,
                                   double factor, void* dst, int dst_stride);
/*
 * Same work on frames already resident in this device's HBM; enqueued on the session stream
 * (or on `stream` if non-NULL: a cudaStream_t), returns without synchronising.
 */
ACB200_API int acb200_process_device(acb200_session* session, const acb200_model* model,
                                     const void* d_src, int w, int h, int c, int src_stride, int elem_type,
                                     double factor, void* d_dst, int dst_stride, void* stream);
ACB200_API int acb200_session_sync(acb200_session* session);

/*
 * One planar / semi-planar YUV video frame -- what every video caller of the reference does per frame
 * (cli/src/Main.cpp:183-206, filter/vapoursynth/src/Filter.cpp:31-42): plane 0 is the 1-channel luma plane and goes through
 * the network (Processor::process), every further plane is chroma (1 channel each: I420/I422/I444, or one interleaved
 * 2-channel plane: NV12/P010...) and goes through the Catmull-Rom resize (ac::core::resize(srcp, dstp, 0.0, 0.0)) to the
 * size of its destination plane.  Here all of it is ONE submission on the session stream.
 *   acb200_plane   same members, in the same order, as one entry of ac::video::Frame::plane
 *                  (video/include/AC/Video/Pipeline.hpp:18-24): a caller can pass `frame.plane` directly
 *   elem_type      ACB200_UINT8 / ACB200_UINT16 (also the float types; `shift` is ignored for them, like ac::core::shl)
 *   shift          for 10/12-bit samples stored LSB-aligned in 16-bit words (cli/src/Main.cpp:175): luma is shifted left by
 *                  `shift` bits before the network and the result shifted right again (ac::core::shl / shr,
 *                  core/src/ImageProcess.cpp:601-616); the source plane itself is not modified; chroma is not shifted
 *   dst planes     caller-allocated; plane 0 must be int(w * factor) x int(h * factor), chroma planes at least the source size
 *   factor         any factor >= 1 (powers of two or not), as for images
 */
typedef struct acb200_plane
{
    int width, height, channel, stride;     /* stride in bytes; 0 = tightly packed */
    unsigned char* data;
} acb200_plane;
ACB200_API int acb200_process_frame_host(acb200_session* session, const acb200_model* model,
                                         const acb200_plane* src, const acb200_plane* dst, int planes,
                                         int elem_type, int shift, double factor);
/* the same on device-resident planes (NVDEC / NVENC surfaces); enqueued on `stream` (NULL: the session stream), no sync */
ACB200_API int acb200_process_frame_device(acb200_session* session, const acb200_model* model,
                                           const acb200_plane* d_src, const acb200_plane* d_dst, int planes,
                                           int elem_type, int shift, double factor, void* stream);

/*
 * Multi-GPU sharding of ONE very large image into halo-overlapped row bands (no collective: every band is independent,
 * the host already holds the whole source).  Band `band` of `n_bands` covers a contiguous range of source rows; the
 * session's GPU receives those rows plus the network's context rows, and only the band's own output rows are copied
 * back into `dst` (the full destination image).  Bands reproduce the whole-image result bit for bit.
 *   acb200_model_halo  rows of input context per 2x pass (= 3x3 layers on the path)
 *   acb200_band_plan   the source rows [src_y0, src_y1) a band reads and the output rows [out_y0, out_y1) it owns
 */
ACB200_API int acb200_model_halo(const acb200_model* model);
ACB200_API int acb200_band_plan(int h, double factor, int halo, int n_bands, int band, int* src_y0, int* src_y1, int* out_y0, int* out_y1);
ACB200_API int acb200_process_host_band(acb200_session* session, const acb200_model* model,
                                        const void* src, int w, int h, int c, int src_stride, int elem_type,
                                        double factor, int n_bands, int band, void* dst, int dst_stride);

/*
 * Stand-alone image ops of the hot path on HOST images (reference: core/src/ImageProcess.cpp:38-61,
 * 113-138,191-215,275-308 and the Catmull-Rom upscale of core/src/ImageResize.cpp:136-272).
 */
ACB200_API int acb200_rgb2yuv_host(acb200_session* session, const void* src, int w, int h, int c, int src_stride, int elem_type,
                                   void* y, int y_stride, void* uv, int uv_stride);
ACB200_API int acb200_yuv2rgb_host(acb200_session* session, const void* y, int y_stride, const void* uv, int uv_stride,
                                   int w, int h, int c, int elem_type, void* dst, int dst_stride);
/* packed 1-plane forms YUV[A] <-> RGB[A] (core/src/ImageProcess.cpp:15-37, 87-112, 166-190, 243-274) */
ACB200_API int acb200_rgb2yuv_packed_host(acb200_session* session, const void* src, int w, int h, int c, int src_stride, int elem_type,
                                          void* yuv, int yuv_stride);
ACB200_API int acb200_yuv2rgb_packed_host(acb200_session* session, const void* yuv, int yuv_stride, int w, int h, int c, int elem_type,
                                          void* dst, int dst_stride);
ACB200_API int acb200_resize_catmull_rom_host(acb200_session* session, const void* src, int w, int h, int c, int src_stride,
                                              int elem_type, void* dst, int ow, int oh, int dst_stride);

/*
 * Multi-GPU frame stream (video frames / filter frontends): frame n is dealt to devices[n mod n_devices]; each device runs
 * `workers_per_device` worker threads with their own session; submit() blocks while that device's bounded queue
 * (`queue_depth` frames) is full; next() hands back finished frames strictly in submission order.  Replaces the worker /
 * ordering core of the reference's video/src/Filter.cpp:33-121; no collective is involved.
 */
typedef struct acb200_stream acb200_stream;
ACB200_API int acb200_frame_owner(long long seq, int n_devices);
ACB200_API int acb200_stream_create(const acb200_model* model, const int* devices, int n_devices, int workers_per_device, int queue_depth,
                                    acb200_stream** out);
ACB200_API int acb200_stream_submit(acb200_stream* stream, const void* src, int w, int h, int c, int src_stride, int elem_type, double factor,
                                    void* dst, int dst_stride, long long* seq_out);
/* planar-frame form of submit (acb200_process_frame_host per frame); the plane arrays are copied, the pixel data is not */
ACB200_API int acb200_stream_submit_frame(acb200_stream* stream, const acb200_plane* src, const acb200_plane* dst, int planes,
                                          int elem_type, int shift, double factor, long long* seq_out);
ACB200_API int acb200_stream_next(acb200_stream* stream, long long* seq_out, int* status_out);
ACB200_API void acb200_stream_destroy(acb200_stream* stream);

/*
 * Page-locked host memory for images that are fed from / delivered to host memory (the host-fed path is PCIe-bound: pageable buffers
 * reach about a quarter of the pinned copy rate).  acb200_host_alloc returns NULL when no CUDA device is usable or pinning fails;
 * the drop-in's ac::core::Image allocates its own storage through a pool of these (ACB200_PINNED_IMAGES=0 turns that off).
 */
ACB200_API void* acb200_host_alloc(size_t bytes);
ACB200_API void acb200_host_free(void* p);
/* number of kernels this library has launched since load (all sessions); for bench.py's gpu_launches */
ACB200_API unsigned long long acb200_launch_count(void);
/* elapsed GPU milliseconds of the most recent process_* call's kernels on this session (CUDA events) */
ACB200_API float acb200_session_last_kernel_ms(acb200_session* session);
/*
 * Luma-network engine:
 *   0  exact   fp32 FFMA in the reference FMA-backend summation order: bit-identical to the reference CPU processor
 *   1  tensor  split-fp16 tensor-core MMA: >= 99.9 % of 8-bit samples identical, <= 1 LSB
 *   2  auto    (default) exact for every 2x pass but the last, tensor for the last -- multi-pass factors (4x, 8x)
 *              keep the 8-bit bar because no rounding difference is fed back into a later pass
 */
ACB200_API int acb200_session_set_engine(acb200_session* session, int engine);
/*
 * implementation of the tensor engine:
 *   0  mma.sync (HMMA, operands via ldmatrix)
 *   1  tcgen05 with the feature maps in shared memory (UTCHMMA, SS form; kept for comparison)
 *   2  (default) tcgen05 with the feature maps resident in TMEM (UTCHMMA, TS form with .ashift): ACNet and ACNet-legacy;
 *      ARNet and the wide families run on 0 / their own kernels
 * Both settings can also be given in the environment for callers that never see a session (the reference-facing
 * ac::core::Processor, the C binding, pyac): ACB200_ENGINE = exact | tensor | auto, ACB200_TENSOR_IMPL = mma | tc5 | tm;
 * acb200_session_create fails with ACB200_EINVAL on any other value.
 */
ACB200_API int acb200_session_set_tensor_impl(acb200_session* session, int impl);
/*
 * Colour handling fused into the luma network's kernels: 0 off, 1 (default) 8-bit RGB, 2 8-bit RGBA as well (bit-identical too, but
 * measured 17 % slower than the separate kernels: the alpha channel runs the chroma machinery a second time); environment:
 * ACB200_FUSE = 0 | 1 | 2.  Where it applies --
 * 8-bit RGB[A], factor exactly 2, the TMEM-resident tensor implementation, models that run as two or more segments -- the RGB -> YUV
 * split (core/src/ImageProcess.cpp:38-61) runs inside the first segment's tile load and the Catmull-Rom chroma resize, its
 * re-quantisation and the YUV -> RGB merge (core/src/processor/Processor.cpp:251-253) run inside the last segment's tail: two launches per
 * frame and no intermediate Y plane.  The result is bit-identical to the separate kernels (on = 0), which every other case uses.
 */
ACB200_API int acb200_session_set_fusion(acb200_session* session, int on);

ACB200_API const char* acb200_error_string(int code);
ACB200_API const char* acb200_version(void);

#ifdef __cplusplus
}
#endif

#endif
