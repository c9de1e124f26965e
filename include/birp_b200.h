/*
 * birp_b200.h -- C ABI of the B200-native line-of-sight integrator.
 *
 * This is the drop-in boundary for the one hot path of Zach-Clare/birp_cpp: the
 * per-pixel integration that the reference runs serially inside
 *     Camera::Render() -> Camera::Integrate()          (reference src/Camera.cpp:63-233)
 * calling Space::GetSample 200x per pixel              (reference include/Space.h:21-22)
 * on a DataCube                                        (reference src/DataCube.cpp:203-277)
 * or on the analytic CMEM field                        (reference src/Cmem.cpp:117-288).
 *
 * Plain pointers and sizes only.  Every entry point returns 0 on success and a
 * BIRP_ERR_* code otherwise; nothing throws across this boundary.  The C++ classes in
 * birp_cpp_b200/include (same names and signatures as the reference's include/ *.h)
 * sit on top of it and turn error codes back into the reference's exceptions.
 *
 * Calls are synchronous on return unless the name ends in _device (those enqueue on
 * the stream given and return).  One host thread per GPU; the library keeps one
 * context (streams, staging buffers, ray tables) per device.
 *
 * There is no CPU fallback: without a CUDA device every compute entry point fails
 * with BIRP_ERR_CUDA.  The host-math entry points (frame setup, ray table, CMEM
 * constants) are per-frame / per-run setup that the reference also does once outside
 * its pixel loop; they run on the host with glibc so their outputs are bit-identical
 * to the reference's.
 */
#ifndef BIRP_B200_H
#define BIRP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    BIRP_OK = 0,
    BIRP_ERR_CUDA = 1,     /* CUDA runtime error, or no device / extension unusable */
    BIRP_ERR_ARG = 2,      /* bad argument (null pointer, non-positive size, ...) */
    BIRP_ERR_SINGULAR = 3, /* rotation matrix has determinant 0: reference throws
                              std::runtime_error("Determinant is 0") (src/Helper.cpp:242-244) */
    BIRP_ERR_RANGE = 4,    /* a sample coordinate was NaN inside the cube bounds test: the
                              reference throws std::out_of_range from .at()
                              (src/DataCube.cpp:245-252,273) */
    BIRP_ERR_NOMEM = 5
};

/* flags for the render entry points */
enum {
    BIRP_NEAREST = 0,        /* DataCube::SetTrilinear(false), the reference default (src/main.cpp:56) */
    BIRP_TRILINEAR = 1,      /* DataCube::SetTrilinear(true) */
    BIRP_ACCUMULATE = 2,     /* start each pixel from the value already in `images`
                                (the reference never clears Camera::image, src/Camera.cpp:223);
                                without it pixels start from 0 and `images` is overwritten */
    BIRP_CMEM_FAST = 4,      /* CMEM only, opt-in: tolerance mode.  The region of a sample is decided like in the default
                                path (region table, fp32 radius estimates with margins); its VALUE is computed in fp32
                                with the hardware approximations (~1e-6 relative) instead of fp64-then-round, and a
                                sample next to a boundary goes through the fp32 math library.  >= 99.9 % of pixels
                                within 1e-5 of the reference; the default path reproduces the reference's samples bit
                                for bit in 99 % of the cases */
    /* diagnostics: force the slower, equally exact kernel variants (see DESIGN.md "variants") */
    BIRP_DIV_TWO_STEP = 8,   /* two residual corrections in the division by the spacing */
    BIRP_DIV_ONE_STEP = 524288, /* one residual correction even where the multiply by the reciprocal alone is exact (a
                                   cube whose spacings are powers of two) */
    BIRP_DIV_IEEE = 16,      /* plain IEEE division by the spacing */
    BIRP_GENERAL_LERP = 32,  /* full std::lerp semantics even for an all-positive cube */
    /* corner fetch of the trilinear kernel; default: one 32-byte load per sample from the oct copy of
       the grid (see BIRP_NO_OCTS), else one 16-byte load per z layer from the quad copy */
    BIRP_NO_TEXTURE = 64,    /* no texture gathers */
    BIRP_TEXTURE_ONLY = 128, /* all eight corners with two texture gathers */
    BIRP_NO_QUADS = 256,     /* do not use the quad copy (lower layer: one texture gather, upper layer:
                                four global loads; with BIRP_NO_TEXTURE: eight global loads) */
    BIRP_QUADS_ONLY = 512,   /* the quad copy (two 16-byte loads) even when the oct copy exists */
    BIRP_CMEM_LIBM64 = 1024, /* CMEM only: the fp64-then-round evaluation through the CUDA fp64 math library
                                instead of the library's own fp64 blocks (same results but for ~1e-8 of the
                                calls; slower; kept as the cross-check of those blocks) */
    BIRP_CMEM_NO_FILTER = 2048, /* CMEM only: evaluate both radii exactly for every sample instead of deciding
                                  the region of samples far from a boundary with fp32 estimates (identical
                                  results; kept as the cross-check of the filter) */
    BIRP_NO_OCTS = 8192,        /* DataCube, trilinear: do not use the oct copy of the grid (the default corner fetch:
                                   ONE 32-byte load per sample from a copy that stores both z layers of every cell
                                   together, 8x the grid in device memory); with this flag, two 16-byte loads from
                                   the quad copy */
    BIRP_LEAN_MARCH = 16384,    /* accepted and ignored (round 1's opt-in; the lean march is now the default where it applies) */
    BIRP_NO_LEAN_MARCH = 65536, /* DataCube, trilinear: keep the general march loop.  By default a positive cube of at
                                   most 2^24 voxels whose value range (found at upload) keeps every scaled sample inside
                                   [2^-120, 2^120) computes the cell index in fp32 and drops the range / midpoint tests of the
                                   scale shortcut, which that range makes redundant (identical results, four instructions
                                   fewer per sample, measured +1.9 %); this flag keeps the general loop as the cross-check */
    BIRP_CMEM_LIBM32 = 262144,   /* CMEM only, diagnostic: the CUDA fp32 math library for every sample, no table, no filter
                                    (round 1's BIRP_CMEM_FAST; slower than the default) */
    BIRP_CMEM_NO_TABLE = 131072, /* CMEM only: do not use the region table of the parameter set (256 x 64 cells over
                                    cos(theta) x cos(phi), each with bounds of both radii over the cell and its neighbours:
                                    a sample far from both boundaries is classified by one 16-byte load instead of the
                                    fp32 radius estimates); identical results, kept as the cross-check */
    BIRP_FAST_MARCH = 32768,    /* DataCube, trilinear, opt-in: tolerance mode.  The default march is bit-identical to the
                                   reference Camera; with this flag the VALUE of an in-bounds sample is computed with fused
                                   arithmetic (one FFMA per coordinate, the cell's interpolant as a polynomial from one
                                   32-byte record, the scale applied once per ray): 28 instead of 60 instructions in the loop.
                                   Which steps are in bounds and where a ray ends are still decided with the reference's own
                                   arithmetic, so images agree with the reference to ~1e-6 relative per pixel (north_star's
                                   tolerance: 1e-5 per pixel, 1e-4 integrated flux).  Needs a further copy of the grid (8x,
                                   built at the first such render; without room for it the exact march runs) */
    BIRP_WINDOW_PER_STEP = 4096 /* DataCube only: test the cube bounds window at every step of the march instead
                                   of locating the end of each ray's in-bounds run once by bisection over the
                                   (monotone) ray table (identical results; kept as the cross-check) */
};

/* Opaque, device-resident emissivity grid (replaces DataCube::slices, include/DataCube.h:42). */
typedef struct birp_cube birp_cube;

/*
 * Per-frame camera state, computed once per frame on the host exactly as the reference
 * does before its pixel loop (src/Camera.cpp:85-137 with Orient :309-417 and
 * PointToPlane :419-508).  64 bytes.
 */
typedef struct birp_frame {
    float pos[3];     /* Camera::position */
    float rinv[9];    /* rotation_inverse, row-major (src/Camera.cpp:135) */
    float tan_half_x; /* tanf(to_rad * (fov_x / 2))   (src/Camera.cpp:150) */
    float tan_half_y; /* tanf(to_rad * (fov_y / 2))   (src/Camera.cpp:151) */
    float aspect;     /* (float)W / (float)H          (src/Camera.cpp:137) */
    float sun_angle;  /* Camera::Orient() result, diagnostic only */
} birp_frame;

/*
 * CMEM state after CMEM::Init (private members of include/Cmem.h:43-75 that GetSample
 * reads, src/Cmem.cpp:117-288).  27 floats.
 */
typedef struct birp_cmem_params {
    float p0, p1, p2, p3;
    float bs, A1, A2, B, alpha, beta;
    float ay_bs, az_bs;
    float r0_lin, charlie, dn, ds, theta_n, theta_s;
    float a5, a21;       /* a[5], a[21] of DefineLinearCoeffs (src/Cmem.cpp:253) */
    float beta_c[4];
    float pressure_dynamic, pressure_magnetic, density;
} birp_cmem_params;

/* ------------------------------------------------------------------ library / device */
const char* birp_version(void);
const char* birp_last_error(void);          /* thread-local message of the last failure */
int birp_device_count(int* count);
int birp_set_device(int ordinal);           /* cudaSetDevice for the calling thread */
int birp_host_alloc(size_t bytes, void** out); /* pinned host memory for images / frames */
int birp_host_free(void* p);
/* page-lock memory the caller already owns (e.g. a POSIX shared-memory segment that several processes drain their
   images into) so that copies to and from it run asynchronously at full PCIe rate; undo with birp_host_unregister */
int birp_host_register(void* p, size_t bytes);
int birp_host_unregister(void* p);

/* ------------------------------------------------------------------ host-side setup */
/* replaces Camera::Integrate's per-frame prologue + Camera::Orient (src/Camera.cpp:85-137,309-508) */
int birp_frame_setup(const float pos[3], float aim_x, float fov_x_deg, float fov_y_deg, int W, int H,
                     birp_frame* out);
/* the same with the whole aim vector of Camera::SetAim(x, y, z): Camera::Orient builds its look direction from all
   three components (src/Camera.cpp:323-327), theta and phi only from aim x (:99-104); birp_frame_setup(pos, aim_x)
   is this call with aim = {aim_x, 0, 0} */
int birp_frame_setup_aim(const float pos[3], const float aim[3], float fov_x_deg, float fov_y_deg, int W, int H,
                         birp_frame* out);
/* the same for n poses (xyz interleaved) with a common aim / field of view: an orbit sweep's frames,
   computed on up to `nthreads` host threads (0 = choose) */
int birp_frame_setup_batch(const float* poses, int n, float aim_x, float fov_x_deg, float fov_y_deg, int W,
                           int H, birp_frame* out, int nthreads);
/* The same on the device, for sweeps whose poses never visit the host: d_poses (xyz interleaved) ->
   d_frames, one thread per pose, enqueued on `stream`.  Same operation order as the host version
   (one shared source, birp_frame_math.h) with every libm call evaluated in fp64 and rounded once,
   so a frame differs from the host's only where a glibc fp32 result is within ~1e-8 of a rounding
   boundary (measured: see tests/test_gpu_parity.py::test_device_frame_setup).  A singular rotation
   sets a status bit: birp_check_device_status returns BIRP_ERR_SINGULAR. */
int birp_frame_setup_device(const float* d_poses, int n, float aim_x, float fov_x_deg, float fov_y_deg, int W, int H,
                            birp_frame* d_frames, void* stream);
/* n camera positions, uniform in time, on a Kepler ellipse with its apogee at `inclination_deg` above
   the +x axis in the x-z plane (the synthetic SMILE-like sweep of SURVEY 8(d) config 5; fp64 Newton
   on Kepler's equation): host and device versions, poses xyz interleaved */
int birp_orbit_poses(int n, double perigee, double apogee, double inclination_deg, float* poses);
int birp_orbit_poses_device(int n, double perigee, double apogee, double inclination_deg, float* d_poses, void* stream);
/* replaces Camera::GenerateRayDistWidth (src/Camera.cpp:69-81); dist_out may be NULL */
int birp_ray_table(int nsamples, int maxdist, float* dist_out, float* ray_width_out);
/* replaces CMEM::Init(bool, ...) (src/Cmem.cpp:38-115); has_v/has_b = 0 mean "empty vector" */
int birp_cmem_init(int has_v, const int v[3], int has_b, const int b[3], float dipole, float p0, int p1,
                   int p2, int p3, float B, float alpha, float beta, float bs, float A1, float A2,
                   float ay_bs, float az_bs, float density, birp_cmem_params* out);

/* ------------------------------------------------------------------ cube residency */
/* zyx: nz*ny*nx floats, x fastest (slices[z][y][x], include/DataCube.h:42).
   origin = coords_{x,y,z}[0], spacing = DataCube::spacing (src/DataCube.cpp:140-191).
   Besides the linear grid the upload builds three verified copies for the trilinear corner fetch -- oct records
   (8x the grid), quads (4x), a layered texture (1x) -- each only when it fits into half of the free device
   memory and passes its self-check; BIRP_DISABLE_OCTS / BIRP_DISABLE_QUADS / BIRP_DISABLE_TEXTURE in the
   environment skip a copy (the kernels fall back to the next fetch form, results are identical). */
int birp_cube_upload(const float* zyx, int nx, int ny, int nz, const float origin[3],
                     const float spacing[3], birp_cube** out);
/* same, the grid already in device memory (e.g. after an NCCL broadcast); copied */
int birp_cube_upload_device(const float* d_zyx, int nx, int ny, int nz, const float origin[3],
                            const float spacing[3], void* stream, birp_cube** out);
void birp_cube_free(birp_cube* cube);
/* a resident cube copied to another GPU of the box (device to device over NVLink, then the same upload-time
   analysis there); *device = the ordinal a cube lives on */
int birp_cube_replicate(const birp_cube* src, int device, birp_cube** out);
int birp_cube_device(const birp_cube* cube, int* device);
/* what the upload-time analysis found: info[0] = division variant proven exact for this cube's
   spacings (3 the multiply by the reciprocal alone: power-of-two spacings; 0 one correction step, 1 two, 2 IEEE only), info[1] = 1 when every voxel is finite and > 0, 2 when
   finite and > 0 or +0 (each class has its own exact form of std::lerp), 0 otherwise,
   info[2] = 1 when the layered-texture copy exists and passed its gather self-check,
   info[3] = 1 when the quad copy (16-byte corner fetch) exists and passed its self-check */
int birp_cube_info(const birp_cube* cube, int info[4]);
/* *present = 1 when the oct copy (32-byte corner fetch) exists and passed its self-check */
int birp_cube_has_oct_records(const birp_cube* cube, int* present);

/* ------------------------------------------------------------------ the hot path, host buffers
 * frames: nframes records (host).  images: nframes*H*W floats (host), row j = image y, top row
 * first, like Camera::image[j][i].  Only rows [row_begin,row_end) are rendered and written
 * (row_end <= 0 means H): the unit of multi-GPU image sharding.
 * nsamples/maxdist = 200/44 reproduce Camera::Render() (src/Camera.cpp:31,65).
 */
int birp_render_cube(const birp_cube* cube, const birp_frame* frames, int nframes, int W, int H,
                     int row_begin, int row_end, int nsamples, int maxdist, int flags, float* images);
int birp_render_cmem(const birp_cmem_params* params, const birp_frame* frames, int nframes, int W, int H,
                     int row_begin, int row_end, int nsamples, int maxdist, int flags, float* images);
/* One large image over several GPUs in INTERLEAVED row bands (reference pixel loops: src/Camera.cpp:140-141): of the
   bands of `band_rows` rows counted from the top, this call renders and writes those with index = phase (mod
   nphases) -- every GPU gets the same mix of cheap rows (rays that leave the cube early) and expensive ones, which
   contiguous bands do not give.  band_rows should be a multiple of 8 (a warp's pixel tile is 4x8 or 8x4).  Rows of
   other phases are neither read nor written. */
int birp_render_cube_rows(const birp_cube* cube, const birp_frame* frames, int nframes, int W, int H, int band_rows,
                          int phase, int nphases, int nsamples, int maxdist, int flags, float* images);
/* A sweep given as camera POSITIONS (xyz interleaved, common aim / field of view): what the reference would do
   with one Camera per pose -- SetPosition, SetAim(aim_x, 0, 0), Render() -- in one call.  The frame records are
   made by birp_frame_setup_batch chunk by chunk while the GPU renders the previous chunk and the copy engine
   drains the one before; frames_out (nframes records, optional, may be NULL) receives them. */
int birp_render_cube_sweep(const birp_cube* cube, const float* poses, int nframes, float aim_x, float fov_x_deg,
                           float fov_y_deg, int W, int H, int row_begin, int row_end, int nsamples, int maxdist,
                           int flags, birp_frame* frames_out, float* images);
int birp_render_cmem_sweep(const birp_cmem_params* params, const float* poses, int nframes, float aim_x,
                           float fov_x_deg, float fov_y_deg, int W, int H, int row_begin, int row_end, int nsamples,
                           int maxdist, int flags, birp_frame* frames_out, float* images);
/* The same with a scatter of the output frames: frame f of the call is written to image slot
   frame_offset + f * frame_stride of `images` (the plain entry points are offset 0, stride 1).  A sweep whose frames
   are interleaved over several processes (frame f of the sweep -> process f mod N) lands in ONE host buffer that the
   processes share (POSIX shared memory, registered with the CUDA driver by each of them): every GPU drains its own
   share over its own PCIe link. */
int birp_render_cube_sweep_strided(const birp_cube* cube, const float* poses, int nframes, float aim_x, float fov_x_deg,
                                   float fov_y_deg, int W, int H, int row_begin, int row_end, int nsamples, int maxdist,
                                   int flags, birp_frame* frames_out, float* images, long long frame_offset,
                                   long long frame_stride);
int birp_render_cmem_sweep_strided(const birp_cmem_params* params, const float* poses, int nframes, float aim_x,
                                   float fov_x_deg, float fov_y_deg, int W, int H, int row_begin, int row_end,
                                   int nsamples, int maxdist, int flags, birp_frame* frames_out, float* images,
                                   long long frame_offset, long long frame_stride);
/* One sweep over several GPUs of the box from ONE process (what north_star calls "frames sharded across the GPUs of
   one box, images gathered"): cubes[d] is the same grid resident on a different device (birp_cube_upload there, or
   birp_cube_replicate), ndevices of them.  One host thread per device; each takes the next chunk of frames from a
   shared counter (the frames of an orbit do not cost the same), makes its frame records on the host while its GPU
   renders the previous chunk, and drains its images straight into `images` (pinned memory: birp_host_alloc) over
   its own PCIe link.  Same results as the single-GPU call, frame for frame.  The first failure stops the hand-out
   and is returned.  The calling thread's current device is restored. */
int birp_render_cube_sweep_multi(const birp_cube* const* cubes, int ndevices, const float* poses, int nframes,
                                 float aim_x, float fov_x_deg, float fov_y_deg, int W, int H, int row_begin, int row_end,
                                 int nsamples, int maxdist, int flags, birp_frame* frames_out, float* images);
int birp_render_cmem_sweep_multi(const birp_cmem_params* params, const int* devices, int ndevices, const float* poses,
                                 int nframes, float aim_x, float fov_x_deg, float fov_y_deg, int W, int H, int row_begin,
                                 int row_end, int nsamples, int maxdist, int flags, birp_frame* frames_out,
                                 float* images);
/* src/Camera.cpp:201-223 on samples[nrays][nsamples] taken on the host from a user-defined Space
   subclass (e.g. the reference tests' CubeFake, tests/Camera_test.cpp:25-29). image: nrays floats. */
int birp_accumulate(const float* samples, long nrays, int nsamples, float ray_width, int flags,
                    float* image);
/* world-space unit ray direction of every pixel of one frame, dirs[H*W*3] (host out): the GPU ray
   generator (src/Camera.cpp:148-168) on its own, for callers that must sample on the host */
int birp_ray_directions(const birp_frame* frame, int W, int H, float* dirs);
/* Space::GetSample on n points, xyz interleaved (host in, host out) */
int birp_sample_cube(const birp_cube* cube, const float* xyz, long n, int flags, float* out);
int birp_sample_cmem(const birp_cmem_params* params, const float* xyz, long n, int flags, float* out);

/* ------------------------------------------------------------------ the hot path, device buffers
 * d_frames / d_images are device pointers; the work is enqueued on `stream` (a cudaStream_t,
 * NULL = default stream) and the call returns without synchronising.
 * d_valid_samples (optional, may be NULL): incremented by the number of samples that were
 * evaluated and accumulated (the live prefix of every ray), for roofline accounting.
 */
int birp_render_cube_device(const birp_cube* cube, const birp_frame* d_frames, int nframes, int W, int H,
                            int row_begin, int row_end, int nsamples, int maxdist, int flags,
                            float* d_images, unsigned long long* d_valid_samples, void* stream);
int birp_render_cmem_device(const birp_cmem_params* params, const birp_frame* d_frames, int nframes,
                            int W, int H, int row_begin, int row_end, int nsamples, int maxdist, int flags,
                            float* d_images, unsigned long long* d_valid_samples, void* stream);
/* The same with a scatter of the output frames: frame f of the call is written to image slot
 * frame_offset + f * frame_stride of d_images (the plain entry points are offset 0, stride 1).
 * d_images may be memory of ANOTHER GPU of the box mapped with birp_ipc_import: the kernel's final
 * pixel stores then travel over NVLink, i.e. a sweep sharded over several processes is gathered on
 * one rank by the render kernel itself, with no separate collective. */
int birp_render_cube_device_strided(const birp_cube* cube, const birp_frame* d_frames, int nframes, int W, int H,
                                    int row_begin, int row_end, int nsamples, int maxdist, int flags,
                                    float* d_images, long long frame_offset, long long frame_stride,
                                    unsigned long long* d_valid_samples, void* stream);
int birp_render_cmem_device_strided(const birp_cmem_params* params, const birp_frame* d_frames, int nframes,
                                    int W, int H, int row_begin, int row_end, int nsamples, int maxdist, int flags,
                                    float* d_images, long long frame_offset, long long frame_stride,
                                    unsigned long long* d_valid_samples, void* stream);
/* the interleaved row bands of birp_render_cube_rows, device buffers (d_images may be another GPU's memory) */
int birp_render_cube_device_rows(const birp_cube* cube, const birp_frame* d_frames, int nframes, int W, int H,
                                 int band_rows, int phase, int nphases, int nsamples, int maxdist, int flags,
                                 float* d_images, long long frame_offset, long long frame_stride,
                                 unsigned long long* d_valid_samples, void* stream);
/* device memory that other processes on the same box can map (CUDA IPC): the owner allocates and exports
   a 64-byte handle, a peer imports it and gets a pointer valid in its own process */
int birp_device_alloc(size_t bytes, void** d_ptr);
int birp_device_free(void* d_ptr);
int birp_ipc_export(const void* d_ptr, unsigned char handle[64]);
int birp_ipc_import(const unsigned char handle[64], void** d_ptr);
int birp_ipc_close(void* d_ptr);

/* Completion flags for results that several processes' kernels store into ONE GPU's memory (birp_*_device_strided with
   an imported pointer): birp_peer_signal enqueues, after the caller's render on `stream`, one store of `value` (the
   step number) into *d_flag -- the caller's slot of a flag array in the gathering GPU's memory, mapped like the
   result buffer; birp_peer_wait makes the gathering GPU's `stream` wait until all n slots hold `value` (or a later
   one).  A 5 us replacement for a one-element all-reduce; the wait gives up after ~2 s and raises the device status
   (birp_check_device_status returns BIRP_ERR_CUDA). */
int birp_peer_signal(unsigned* d_flag, unsigned value, void* stream);
int birp_peer_wait(const unsigned* d_flags, int n, unsigned value, void* stream);

/* device error word of the calling thread's device (BIRP_ERR_RANGE / BIRP_ERR_SINGULAR conditions raised
   by *_device launches); synchronises `stream`, returns BIRP_OK or the error and clears it */
int birp_check_device_status(void* stream);

/* diagnostics: max error of the library's own fp64 exp2 / log2 / pow / exp (relative), acos and the
   angle bundle sin, cos, cos(a/2) (absolute) -- the blocks of the CMEM kernel -- against the CUDA fp64
   math library on n pseudo-random arguments */
int birp_selftest_cmem_math(long n, double max_err[6]);

/* diagnostics: the CMEM kernel takes r / 10 (src/Cmem.cpp:158,165) as a multiplication by RN(1/10) with one exact
   correction step where r is in [2^-100, 2^100), and as the division elsewhere.  Compares that with the IEEE division
   for EVERY fp32 bit pattern on the device: mismatches[0] = patterns where the kernel's form differs (must be 0),
   mismatches[1] = positive normal r where the multiplication form alone would (419 431: the subnormal quotients) */
int birp_selftest_div10(unsigned long long mismatches[2]);

/* diagnostics of the CMEM region filter: out[0] / out[1] = the largest fraction of the filter's margin that
   the fp32 estimate of the magnetopause / bow-shock radius was off by, over n pseudo-random sample positions
   (must stay far below 1), out[2] = fraction of the positions whose direction the filter trusts */
int birp_selftest_cmem_filter(const birp_cmem_params* params, long n, double out[3]);

/* diagnostics of the CMEM region table of a parameter set (built at first use on the calling thread's device): the
   march's sampler with the table against the unfiltered evaluation on n pseudo-random positions; out[0] = samples that
   differ bitwise (must be 0), out[1] = fraction of the positions the table decided, out[2] = 1 when the table is in use
   (0: it failed its self-test at first use, or BIRP_DISABLE_CMEM_TABLE is set) */
int birp_selftest_cmem_table(const birp_cmem_params* params, long n, double out[3]);

/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long birp_launch_count(void);

/* diagnostics: launch-shape knobs of the calling thread's device, for A/B measurements and for the tests that run
   every code variant (results never depend on them beyond what the flags document): "tile_w_log2" (-1 = choose; a
   warp's 32 rays form a 2^v x 2^(5-v) pixel tile), "tile_swizzle" (1: four consecutive lanes own a 2x2 pixel block),
   "fast_variant" (-1 = default; 0..4: how the tolerance-mode march rounds its cell index, see fast_sample),
   "fast_kstep" (0: the tolerance-mode march reads the ray table instead of computing step distances), "fast_coop" (1:
   the warp-per-ray experiment), "cta_warps" (0 = choose: 4 warps per CTA for frames of 2^19 rays and more, else 8;
   2 | 4 | 8 force it).  Most of these
   knobs are also read from the environment (BIRP_TILE_W_LOG2, BIRP_TILE_SWIZZLE, BIRP_FAST_VARIANT, BIRP_FAST_KSTEP,
   BIRP_CTA_WARPS)
   when the device context is created. */
int birp_set_tuning(const char* name, int value);

#ifdef __cplusplus
}
#endif
#endif /* BIRP_B200_H */
