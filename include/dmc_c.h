/* dmc_c.h -- C ABI of libdmc_b200.so: the B200 (sm_100a) implementation of the post filter set for decoded
 * depth maps of Wavelet303/DepthMapCompression.
 *
 * The reference has no FFI layer: its boundary is the C++ header PostFilterSetForDepthCoding/filter.h (plus the
 * converter declarations of util.h that the chain uses).  include/filter.h in this repo keeps those declarations
 * byte-for-byte and forwards every call to the entry points below; each entry point names the reference
 * declaration it replaces (paths relative to /root/reference/PostFilterSetForDepthCoding/).
 *
 * Conventions
 *   - plain pointers and sizes only; no exceptions cross this boundary; every call returns a dmc_status.
 *   - an image is {data, rows, cols, cvtype, step, mem}; cvtype is OpenCV's encoding depth + ((channels-1) << 3)
 *     with depth 0 = 8U, 2 = 16U, 3 = 16S, 5 = 32F, 6 = 64F; step is the row pitch in bytes (0 = dense);
 *     mem says whether `data` is host or device memory.  Host images are staged through device memory by the
 *     library (H2D, kernels, D2H) and the call returns when `dst` is valid, like the reference's synchronous API.
 *     Device images are processed in stream order on the context's stream; call dmc_synchronize() (or use your own
 *     stream via dmc_set_stream) before reading them.
 *   - src and dst may alias (in-place is legal for every operator, as in the reference).
 *   - the library never allocates caller-visible memory: dst must be allocated by the caller (include/filter.h does
 *     the `dst.create()` the reference does).
 *   - DMC_UNSUPPORTED mirrors the reference's silent no-ops (e.g. 8U + FULL_KERNEL_PAIR): dst is left untouched.
 *   - one dmc_ctx per (host thread, device); a context is not thread-safe (neither is a PostFilterSet instance).
 *   - there is NO CPU fallback: without a CUDA device dmc_create fails with DMC_ERR_CUDA.
 */
#ifndef DMC_C_H
#define DMC_C_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    DMC_OK = 0,
    DMC_UNSUPPORTED = 1,   /* reference's silent no-op for this (type, method) pair; dst untouched */
    DMC_ERR_TYPE = -1,     /* CV_Assert on type would have fired in the reference */
    DMC_ERR_SIZE = -2,     /* size / step / null-pointer problem */
    DMC_ERR_CUDA = -3,     /* CUDA runtime error (see dmc_last_error) */
    DMC_ERR_ARG = -4       /* invalid parameter (even median size, radius out of range, ...) */
} dmc_status;

enum { DMC_MEM_HOST = 0, DMC_MEM_DEVICE = 1 };
enum { DMC_8U = 0, DMC_16U = 2, DMC_16S = 3, DMC_32F = 5, DMC_64F = 6 };
/* filter.h:23-28 */
enum { DMC_FULL_KERNEL = 0, DMC_FULL_KERNEL_PAIR = 1, DMC_SEPARABLE_KERNEL = 2 };
/* util.h:19-23 */
enum { DMC_FILL_DISPARITY = 0, DMC_FILL_DEPTH = 1 };
/* cv::BORDER_REPLICATE: the only border the reference's range filter is ever called with (filter.h:29 default) */
enum { DMC_BORDER_REPLICATE = 1 };

#define DMC_MAKETYPE(depth, cn) (((depth) & 7) + (((cn) - 1) << 3))
#define DMC_MAX_RADIUS 10   /* trackbar maxima main.cpp:87-93 */

typedef struct dmc_image {
    void* data;
    int rows, cols;
    int cvtype;
    size_t step;   /* bytes per row; 0 = cols * elemSize */
    int mem;       /* DMC_MEM_HOST or DMC_MEM_DEVICE */
} dmc_image;

typedef struct dmc_ctx dmc_ctx;

/* ---- context -------------------------------------------------------------------------------------------- */
int dmc_device_count(void);
int dmc_create(int device, dmc_ctx** out);
void dmc_destroy(dmc_ctx* ctx);
const char* dmc_last_error(const dmc_ctx* ctx);           /* valid until the next call on ctx; ctx may be NULL */
int dmc_set_stream(dmc_ctx* ctx, void* cuda_stream);      /* use the caller's cudaStream_t (NULL = own stream) */
void* dmc_get_stream(dmc_ctx* ctx);
int dmc_synchronize(dmc_ctx* ctx);
uint64_t dmc_kernel_launches(const dmc_ctx* ctx);         /* kernels launched by this context so far */
void* dmc_host_alloc(size_t bytes);                       /* pinned host memory for the streaming entry points */
void dmc_host_free(void* p);
/* Pins and device-maps memory the caller already owns (a cv::Mat's buffer, a numpy array): cudaHostRegister.  Host images of
 * up to 2 MB that lie in pinned memory (this call, dmc_host_alloc, cudaMallocHost) are processed in place over the host link
 * by the single-image entry points -- no staging copies (640x480 filterDisp8U2Depth32F host to host: 152 -> about 50 us).
 * Unregister before the memory is freed. */
int dmc_host_register(void* p, size_t bytes);
int dmc_host_unregister(void* p);
int dmc_version(void);
/* Device-resident frame batches: number of frame groups in flight on separate streams (1..3, default 1). */
int dmc_set_lanes(dmc_ctx* ctx, int lanes);

/* Optional CUDA-event timing of the chain's stages (used by bench.py for the live roofline figure): while a stage's
 * bit is set in stage_mask, every launch of that stage inside the chain entry points is bracketed by an event pair
 * on the launching stream.  dmc_profile_read synchronises, folds the finished pairs into per-stage totals and
 * returns them (pixels = frames x rows x cols the launches covered). */
enum { DMC_STAGE_MEDIAN = 0, DMC_STAGE_GAUSS = 1, DMC_STAGE_MINMAX = 2, DMC_STAGE_RANGE = 3, DMC_STAGE_COUNT = 4 };
int dmc_profile_enable(dmc_ctx* ctx, int stage_mask);
int dmc_profile_read(dmc_ctx* ctx, int stage, double* total_ms, uint64_t* launches, uint64_t* pixels, int reset);

/* ---- PostFilterSet (filter.h:32-42, postFilterSet.cpp:21-63) --------------------------------------------- */
/* PostFilterSet::operator() filter.h:41: 8UC1 -> 8UC1 */
int dmc_post_filter_set(dmc_ctx* ctx, const dmc_image* src, dmc_image* dst, int median_r, int gaussian_r,
                        int minmax_r, int brange_r, int brange_th, int brange_method);
/* PostFilterSet::filterDisp8U2Depth32F filter.h:38: 8UC1 -> 32FC1 */
int dmc_filter_disp8u_depth32f(dmc_ctx* ctx, const dmc_image* src, dmc_image* dst, double focus, double baseline,
                               double amp, int median_r, int gaussian_r, int minmax_r, int brange_r, float brange_th,
                               int brange_method);
/* PostFilterSet::filterDisp8U2Depth16U filter.h:39: 8UC1 -> 16UC1 */
int dmc_filter_disp8u_depth16u(dmc_ctx* ctx, const dmc_image* src, dmc_image* dst, double focus, double baseline,
                               double amp, int median_r, int gaussian_r, int minmax_r, int brange_r, float brange_th,
                               int brange_method);
/* PostFilterSet::filterDisp8U2Disp32F filter.h:40: 8UC1 -> 16UC1 (sic) */
int dmc_filter_disp8u_disp32f(dmc_ctx* ctx, const dmc_image* src, dmc_image* dst, int median_r, int gaussian_r,
                              int minmax_r, int brange_r, float brange_th, int brange_method);

/* Frame batches (video / multi-view): n_frames dense rows x cols frames, back to back, in host or device memory.
 * chain: 0 = operator() (u8 out), 1 = Depth32F (f32 out), 2 = Depth16U (u16 out), 3 = Disp32F (u16 out).
 * With host memory the frames are streamed through the device in chunks over several CUDA streams (H2D of chunk
 * i+1, kernels of chunk i and D2H of chunk i-1 overlap); pinned memory (dmc_host_alloc) is needed for overlap. */
enum { DMC_CHAIN_DISP8U = 0, DMC_CHAIN_DEPTH32F = 1, DMC_CHAIN_DEPTH16U = 2, DMC_CHAIN_DISP32F = 3 };
typedef struct dmc_chain_params {
    int chain;
    int median_r, gaussian_r, minmax_r, brange_r;
    float brange_th;           /* operator(): integer threshold, passed as float like postFilterSet.cpp:62 */
    int brange_method;
    double focus, baseline, amp;   /* depth chains only */
} dmc_chain_params;
int dmc_chain_batch(dmc_ctx* ctx, const void* src, void* dst, int n_frames, int rows, int cols,
                    const dmc_chain_params* p, int mem);
/* The same chain on a batch given as arrays of image descriptors: n_frames CV_8UC1 images of one size, each anywhere in
 * host or device memory with its own row step (e.g. a std::vector<cv::Mat> of video frames or camera views); dsts[i] has
 * the chain's output type.  Frames are packed into dense device buffers chunk by chunk and streamed through the same
 * H2D / kernel / D2H pipeline.  Returns when every dsts[i] is valid. */
int dmc_chain_batch_images(dmc_ctx* ctx, const dmc_image* srcs, dmc_image* dsts, int n_frames, const dmc_chain_params* p);
/* Frame-batch scheduler across the GPUs of one box, in one process (the reference scales by row-striping one image
 * over CPU threads -- cv::parallel_for_, binalyWeightedRangeFilter.cpp:1080; here whole frames go to whole GPUs):
 * one context per listed device, kept between runs; every run cuts the batch into contiguous shards
 * (dmc_shard_frames) and streams each shard through its device on its own host thread, as dmc_chain_batch does.
 * src/dst are HOST buffers (pinned for full overlap).  No data-path exchange between devices: frames are independent. */
typedef struct dmc_sched dmc_sched;
int dmc_sched_create(const int* devices, int n_devices, dmc_sched** out);
void dmc_sched_destroy(dmc_sched* sched);
int dmc_sched_device_count(const dmc_sched* sched);
const char* dmc_sched_last_error(const dmc_sched* sched);
int dmc_sched_chain_batch(dmc_sched* sched, const void* src, void* dst, int n_frames, int rows, int cols, const dmc_chain_params* p);
/* one-shot convenience: create, run, destroy; on error the message is copied into err (may be NULL) */
int dmc_multi_chain_batch(const int* devices, int n_devices, const void* src, void* dst, int n_frames, int rows, int cols,
                          const dmc_chain_params* p, char* err, size_t err_len);
/* Host-link topology.  Every frame of a host-resident batch crosses the host link twice, and on a multi-GPU box the
 * link is not uniform (profiles/r02_hostlink.json: one group of four B200s shares ~51 GB/s each way, the other four reach
 * ~94 GB/s together, all eight together only ~64 GB/s).  dmc_hostlink_probe copies 64 MB blocks both ways on all listed
 * devices at once (~0.1 s, call it while the devices are otherwise idle), then on the subset with the best per-device rate
 * alone, and proposes a routing: gateway[i] is the device whose link should carry device[i]'s host traffic (itself when
 * its own link is as good as any).  dmc_set_gateway makes the host-memory batch entry points of a context stage through
 * the gateway's HBM: H2D / D2H run on the gateway's copy engines, the context's kernels read their input from and write
 * their output to the gateway's memory directly over NVLink / NVSwitch (peer access), no extra copy.  Results are
 * unchanged (bit-identical); -1 restores the context's own link.  dmc_sched_create probes and routes by itself
 * (DMC_NO_TOPOLOGY=1 disables that). */
#define DMC_MAX_DEVICES 16
typedef struct dmc_hostlink_info {
    int n_devices;
    int device[DMC_MAX_DEVICES];
    int gateway[DMC_MAX_DEVICES];        /* proposed routing */
    double loaded_gbs[DMC_MAX_DEVICES];  /* GB/s each way of device[i] while ALL listed devices copy both ways at once */
    double all_gbs;                      /* aggregate GB/s each way, all listed devices on their own links */
    double best_gbs;                     /* aggregate GB/s each way of the proposed link set (== all_gbs if nothing is re-routed) */
    int n_link;                          /* devices whose links carry traffic under the proposed routing */
} dmc_hostlink_info;
int dmc_hostlink_probe(const int* devices, int n_devices, dmc_hostlink_info* info);
/* cudaDeviceReset of a device this process only probed (no live allocations / contexts of the caller on it); the calling
 * thread's current device is `device` afterwards */
int dmc_release_device(int device);
int dmc_set_gateway(dmc_ctx* ctx, int gateway_device);
int dmc_get_gateway(const dmc_ctx* ctx);
/* routing chosen by the scheduler at creation: gateways[i] for the i-th device, -1 = own link */
int dmc_sched_get_routing(const dmc_sched* sched, int* gateways, double* all_gbs, double* best_gbs);

/* Frame-parallel sharding of a batch over `world` ranks (one process per GPU, no collective): frames
 * [*begin, *begin + *count) belong to `rank` (contiguous blocks, sizes differ by at most one). */
int dmc_shard_frames(int n_frames, int rank, int world, int* begin, int* count);

/* ---- decode feeding the chain ------------------------------------------------------------ */
/* Frame-parallel baseline JPEG decode (SOF0, 8-bit, single component), bit-identical to libjpeg(-turbo)'s default
 * JDCT_ISLOW decoder, i.e. to the reference's imdecode(buf, 0) (main.cpp:284, :521) and jpeg_decode()
 * (jpegTurboDemo.cpp:217-271).  `blob` is HOST memory holding the n_frames bitstreams; stream i occupies
 * [offsets[i], offsets[i+1]).  All frames must be rows x cols.  dst receives n_frames dense 8UC1 frames
 * (dst_mem: host or device; with device memory the chain entry points can consume it without leaving the GPU). */
int dmc_jpeg_decode_gray_batch(dmc_ctx* ctx, const void* blob, const uint64_t* offsets, int n_frames, int rows, int cols,
                               void* dst, int dst_mem);
/* Streamed bitstream -> chain (the reference's pointcloudTest loop main.cpp:276-303 for a whole batch): JPEG streams in
 * HOST memory (pinned for full overlap) are copied to the device (about 1/30 of the bytes of the decoded frames), decoded
 * there and run through the chain of *p without leaving the device; dst receives n_frames frames of the chain's output type
 * (dst_mem: host -- pinned for overlap -- or device).  Chunks of frames flow through the same 4-slot H2D / kernels / D2H
 * pipeline as dmc_chain_batch, with no host synchronisation between chunks.  Returns when dst is valid (host) or when the
 * work is queued (device: dmc_synchronize before reading). */
int dmc_chain_batch_jpeg(dmc_ctx* ctx, const void* blob, const uint64_t* offsets, int n_frames, int rows, int cols,
                         void* dst, int dst_mem, const dmc_chain_params* p);
/* Host-only: parses the headers of one stream, reports its size and whether this library can decode it (DMC_OK) or why
 * not (DMC_ERR_TYPE + message in err, may be NULL).  Needs no context and no GPU.  Malformed, truncated or hostile
 * streams (bad segment lengths, Huffman tables that are not prefix codes, ...) are refused here, never crashed on. */
int dmc_jpeg_probe(const void* stream, size_t len, int* rows, int* cols, char* err, size_t err_len);

/* ---- stand-alone operators of filter.h ------------------------------------------------------------------- */
/* binalyWeightedRangeFilter filter.h:29 (binalyWeightedRangeFilter.cpp:1106): 8U/16S/16U/32F x C1/C3 */
int dmc_bwrf(dmc_ctx* ctx, const dmc_image* src, dmc_image* dst, int kernel_w, int kernel_h, float threshold,
             int method, int border_type);
/* Joint (guided) binary-weighted range filter -- an EXTENSION with no counterpart in the reference:
 * src (CV_8UC1) is averaged over the window of binalyWeightedRangeFilter with the binary weights computed on `guide`
 * (CV_8UC3: saturated L1 colour distance of binalyWeightedRangeFilter.cpp:297-301; CV_8UC1: absolute difference) instead
 * of on src itself.  Same window, border (REPLICATE), (uchar)threshold, division and rounding as the 8UC1 filter; with
 * guide == src the result equals dmc_bwrf.  method must be DMC_FULL_KERNEL. */
int dmc_joint_bwrf(dmc_ctx* ctx, const dmc_image* src, const dmc_image* guide, dmc_image* dst, int kernel_w, int kernel_h,
                   float threshold, int method);
/* blurRemoveMinMax filter.h:19 (minmaxFilter.cpp:176), blurRemoveMinMaxBase filter.h:20 (:216): same result */
int dmc_blur_remove_minmax(dmc_ctx* ctx, const dmc_image* src, dmc_image* dst, int r);
/* maxFilter / minFilter filter.h:17-18 (minmaxFilter.cpp:314, :394): single channel 8U/16S/16U/32F */
int dmc_max_filter(dmc_ctx* ctx, const dmc_image* src, dmc_image* dst, int kernel_w, int kernel_h, int border_type);
int dmc_min_filter(dmc_ctx* ctx, const dmc_image* src, dmc_image* dst, int kernel_w, int kernel_h, int border_type);
/* boundaryReconstructionFilter filter.h:45 (boundaryReconstructionFilter.cpp:133): single channel, 5 depths */
int dmc_boundary_reconstruction(dmc_ctx* ctx, const dmc_image* src, dmc_image* dst, int kernel_w, int kernel_h,
                                float frec, float color, float space);
/* Fused min-max -> boundary reconstruction -- an EXTENSION (BASELINE north_star): the result of
 * dmc_blur_remove_minmax(src, tmp, minmax_r) followed by dmc_boundary_reconstruction(tmp, dst, ...) in ONE kernel: the
 * min-max image is produced tile by tile in shared memory and never written to HBM, so src is read once.  Single channel
 * 8U / 16U / 16S (the integer depths of minmaxFilter.cpp:48-174); bit-identical to the two calls. */
int dmc_minmax_boundary_reconstruction(dmc_ctx* ctx, const dmc_image* src, dmc_image* dst, int minmax_r, int kernel_w, int kernel_h,
                                       float frec, float color, float space);
/* smallGaussianBlur filter.h:14 (postFilterSet.cpp:4-16): 8UC1 */
int dmc_small_gaussian(dmc_ctx* ctx, const dmc_image* src, dmc_image* dst, int d, double sigma);
/* cv::medianBlur as called at postFilterSet.cpp:23,36,47,59: 8UC1, odd ksize */
int dmc_median_blur(dmc_ctx* ctx, const dmc_image* src, dmc_image* dst, int ksize);

/* ---- depthmapUtil.cpp helpers on the filter path (util.h:16-28) ------------------------------------------- */
int dmc_disp8u2depth32f(dmc_ctx* ctx, const dmc_image* src, dmc_image* dst, float focal_baseline, float a, float b);   /* util.h:28 */
int dmc_depth32f2disp8u(dmc_ctx* ctx, const dmc_image* src, dmc_image* dst, float focal_baseline, float a, float b);   /* util.h:25 */
int dmc_depth16u2disp8u(dmc_ctx* ctx, const dmc_image* src, dmc_image* dst, float focal_baseline, float a, float b);   /* util.h:27 */
int dmc_disp16s2depth16u(dmc_ctx* ctx, const dmc_image* src, dmc_image* dst, float focal_baseline, float a, float b);  /* util.h:26 */
int dmc_fill_occlusion(dmc_ctx* ctx, dmc_image* img, int invalid_value, int disp_or_depth);                           /* util.h:24 */
/* cv::transpose as used between the two fillOcclusion passes of pointcloudTest (main.cpp:258, :260); single channel */
int dmc_transpose(dmc_ctx* ctx, const dmc_image* src, dmc_image* dst);
/* reprojectXYZ(depth, xyz, f) util.h:11: xyz is (rows*cols) x 1 32FC3, dense */
int dmc_reproject_xyz(dmc_ctx* ctx, const dmc_image* depth, dmc_image* xyz, double f);

/* splitBGRLineInterleave filter.h:12 (split.cpp:167-177): 8UC3 / 32FC3 -> single channel 3*rows x cols (a B, a G and an R
 * row per image row: the layout the reference's SSE range filter works on); other types are a silent no-op */
int dmc_split_bgr_line_interleave(dmc_ctx* ctx, const dmc_image* src, dmc_image* dest);

/* ---- point-cloud render (util.h:12-13, :25, :33) ------------------------------------------- */
/* R (3x3), t (3), K (3x3) are row-major doubles (the reference asserts CV_64F, depthmapUtil.cpp:290-294).
 * flags: DMC_RENDER_EXACT_DIVIDE replaces the reference's _mm_rcp_ps (depthmapUtil.cpp:78, emulated from Intel's table)
 * by a true division for every point (the reference's own scalar twin myProjectPoint_BF :99-146). */
enum { DMC_RENDER_EXACT_DIVIDE = 1 };
/* projectPointsSimple util.h:33 (depthmapUtil.cpp:148): xyz 32FC3 (n x 1, dense) -> pt 32FC2 (n x 1, dense) */
int dmc_project_points(dmc_ctx* ctx, const dmc_image* xyz, const double* R, const double* t, const double* K, dmc_image* pt, int flags);
/* projectImagefromXYZ util.h:12-13 (depthmapUtil.cpp:285-448): z-buffer splat of `image` (8UC3) at the projected positions
 * of xyz (32FC3, one point per pixel, dense) into destimage (8UC3, cleared first); depth (32FC1) and pt (32FC2) are the
 * optional outputs of the second overload (NULL to skip).  The reference's SERIAL raster-order semantics (a point tries its
 * neighbour pixels only if it won its own pixel at that moment; isSub quirks included) are reproduced exactly, see
 * csrc/dmc_render.cu.  Synchronous: returns when the outputs are complete. */
int dmc_project_image_from_xyz(dmc_ctx* ctx, const dmc_image* image, dmc_image* destimage, const dmc_image* xyz, const double* R,
                               const double* t, const double* K, int is_sub, dmc_image* depth, dmc_image* pt, int flags);
/* fillSmallHole util.h:25 (depthmapUtil.cpp:187-283): 8UC3; only hole pixels (green == 0) of the interior are written,
 * dest keeps everything else (call it in place, like main.cpp:355, or with dest pre-filled). */
int dmc_fill_small_hole(dmc_ctx* ctx, const dmc_image* src, dmc_image* dest);

#ifdef __cplusplus
}
#endif
#endif
