/*
 * sfb200.h — C ABI of libsfb200.so, the B200-native backend of ShaderFlow's offline render hot loop.
 *
 * The reference (BrokenSource/ShaderFlow v0.11.3) has no FFI of its own: its native back-ends are
 * third-party wheels reached through Python objects (moderngl, numpy.fft, scipy.sparse, turbopipe).
 * Every entry point below names the reference call site(s) it replaces (paths relative to the reference
 * root). The binding a maintainer adds on the reference side is a ctypes stub: see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success or a negative SFB_E* code; sfb_last_error() gives the
 *     thread-local message (reference convention: Python exceptions — the ctypes wrapper raises
 *     RuntimeError, mirroring shader.py:354-355 / exporting.py:157-162 / texture.py:251-252);
 *   - one sfb_ctx per (process, device); all calls for a ctx come from one host thread; work is enqueued
 *     on the ctx stream and is asynchronous unless stated (the GL context had the same rule);
 *   - pointers named *_dev are device pointers (e.g. torch tensor.data_ptr()), *_host are host pointers;
 *   - images are bottom-row-first like GL framebuffers / fbo.read_into (exporting.py:165-174);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with SFB_ECUDA.
 */
#ifndef SFB200_H
#define SFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFB_VERSION 100  /* 0.1.0 */

enum {
    SFB_OK = 0,
    SFB_EINVAL = -1,     /* bad argument */
    SFB_ECUDA = -2,      /* CUDA runtime / driver error, or no device */
    SFB_ENOTFOUND = -3,  /* unknown scene / fragment hash */
    SFB_ENOMEM = -4,
    SFB_EIO = -5,        /* sink write failed */
    SFB_ESTATE = -6,     /* call made in the wrong state */
};

typedef struct sfb_ctx sfb_ctx;
typedef struct sfb_tex sfb_tex;
typedef struct sfb_pipe sfb_pipe;

/* ------------------------------------------------------------------------------------------------ */
/* Library / context — replaces scene.py:143-157 (window + moderngl.Context creation) and
 * scene.py:100-107 (teardown). `stream` is a cudaStream_t (0 = legacy default stream); pass
 * torch.cuda.current_stream().cuda_stream so torch events bracket the work. */
int         sfb_version(void);
const char* sfb_last_error(void);
int         sfb_device_count(int* count);
int         sfb_ctx_create(int device, void* stream, sfb_ctx** out);
int         sfb_ctx_destroy(sfb_ctx* ctx);
int         sfb_ctx_set_stream(sfb_ctx* ctx, void* stream);
int         sfb_sync(sfb_ctx* ctx);                       /* blocks: cudaStreamSynchronize */
/* Lets kernels of this ctx store into memory of `peer_device` (cudaDeviceEnablePeerAccess): the sharded
 * export renders frames straight into rank 0's HBM over NVLink (no reference counterpart: the reference
 * is single-GPU). Idempotent. */
int         sfb_ctx_enable_peer(sfb_ctx* ctx, int peer_device);
/* Kernel launches issued through this ctx since creation (bench.py's `gpu_launches`) */
int         sfb_launch_count(sfb_ctx* ctx, uint64_t* count);

/* ------------------------------------------------------------------------------------------------ */
/* Host-side time base (no GPU). Replaces the freewheel SchedulerTask (scheduler.py:86-89,134-173) +
 * ShaderScene.next's time integration (scene.py:476-479) + BrokenAudioReader.stream's chunk rule
 * (ffmpeg.py:1306-1330). For frame k writes the values modules see DURING that frame:
 *   time[k], dt[k]  — scene.time / scene.dt;   tell[k] — samples consumed after ShaderAudio.update.
 * total_samples < 0 means an endless stream. Any output pointer may be NULL. */
int sfb_frame_clock(int n_frames, double fps, double speed, int samplerate, int channels,
                    int64_t total_samples, double* time_host, double* dt_host, int64_t* tell_host);

/* ------------------------------------------------------------------------------------------------ */
/* Audio: STFT → filterbank spectrogram (kernel K1).
 * Replaces, for a batch of frames at once, BrokenSpectrogram.fft + .next (audio/spectrogram.py:155-176):
 * numpy.fft.rfft (pocketfft, float64) and scipy.sparse.csr_matrix.dot, plus the ring slicing of
 * BrokenAudio.get_last_n_samples (audio/module.py:137-138).
 *
 *   pcm_dev      [channels][n_samples] float32 planar — the whole clip resident in HBM
 *   tell_dev     [n_frames] int64 — samples consumed at each frame (sfb_frame_clock); the window of frame
 *                k is clip samples [tell-N-1, tell-1), zero where negative (newest sample excluded)
 *   fft_n        N = 2^fft_n, 8 <= fft_n <= 13; channels must be 1 or 2
 *   window_kind  SFB_WINDOW_*   (spectrogram.py:90-108)      magnitude_kind SFB_MAGNITUDE_* (:20-26)
 *   csr_*        filterbank (bins x (N/2+1)) in CSR, float32 (spectrogram.py:194-224), device pointers
 *   volume_kind  SFB_VOLUME_* epilogue (:28-41; the reference never applies it — LINEAR = parity)
 *   mag_out_dev  optional [n_frames][channels][N/2+1] float32 (NULL to skip)
 *   spec_out_dev [n_frames][bins][channels] float32 — texel b of frame k = (L_b, R_b), the layout
 *                ShaderSpectrogram writes to its RG32F texture column (spectrogram.py:306-311)
 */
enum { SFB_WINDOW_HANNING = 0, SFB_WINDOW_HANN_POISSON = 1, SFB_WINDOW_NONE = 2 };
enum { SFB_MAGNITUDE_POWER = 0, SFB_MAGNITUDE_AMPLITUDE = 1 };
enum { SFB_VOLUME_LINEAR = 0, SFB_VOLUME_SQRT = 1, SFB_VOLUME_DBFS = 2, SFB_VOLUME_DBFS_TREMX = 3 };

int sfb_stft_mel(sfb_ctx* ctx, const float* pcm_dev, int64_t n_samples, int channels, int fft_n,
                 const int64_t* tell_dev, int n_frames, int window_kind, int magnitude_kind,
                 const int32_t* csr_indptr_dev, const int32_t* csr_indices_dev, const float* csr_data_dev,
                 int bins, int volume_kind, float* mag_out_dev, float* spec_out_dev);

/* Audio: everything else the audio modules compute per frame (kernel K2), for a batch of frames.
 *   (1) second-order dynamics over the spectrogram columns — DynamicNumber.next (dynamics.py:197-250)
 *       as ShaderSpectrogram.update drives it (spectrogram.py:303-311): float32 state, in place on
 *       spec_inout_dev [n_frames][bins][channels]; a recurrence over frames, run as one sequential scan;
 *   (2) volume / std targets over the last 0.1 s (audio/module.py:457-458) and their two ShaderDynamics
 *       (audio/module.py:413-421; float64 state): scalars_out_dev [n_frames][SFB_SCALARS] float64 =
 *       {volume, volume_integral, std, volume_target, std_target};
 *   (3) waveform reducer rows (audio/waveform.py:14-22,64-87): wave_out_dev [n_frames][points][channels].
 * dt_dev [n_frames] float64 is scene.dt per frame (sfb_frame_clock). Any of spec_inout_dev /
 * scalars_out_dev / wave_out_dev may be NULL to skip that part.
 * Synchronisation: with wave_out_dev the call reads tell[0] and tell[n_frames-1] back to size its chunk table, i.e. it
 * blocks until everything enqueued on the ctx stream so far has finished (once per export, not per frame); without
 * it the call only enqueues.
 */
enum { SFB_SCALAR_VOLUME = 0, SFB_SCALAR_VOLUME_INTEGRAL = 1, SFB_SCALAR_STD = 2,
       SFB_SCALAR_VOLUME_TARGET = 3, SFB_SCALAR_STD_TARGET = 4, SFB_SCALARS = 5 };
enum { SFB_REDUCER_AVERAGE = 0, SFB_REDUCER_RMS = 1, SFB_REDUCER_STD = 2 };

typedef struct sfb_dynamics_params {   /* dynamics.py:140-158 */
    double frequency, zeta, response, precision;
} sfb_dynamics_params;

int sfb_audio_track(sfb_ctx* ctx, const float* pcm_dev, int64_t n_samples, int channels, int samplerate,
                    const int64_t* tell_dev, const double* dt_dev, int n_frames,
                    float* spec_inout_dev, int bins, const sfb_dynamics_params* spec_dynamics,
                    double* scalars_out_dev,
                    float* wave_out_dev, int wave_points, int wave_chunk, int wave_reducer);

/* Audio file ingestion without ffmpeg. The reference decodes files through an ffmpeg child asked
 * for interleaved float32 (BrokenAudioReader.stream, ffmpeg.py:1279-1330) and transposes every chunk into its host
 * ring (audio/module.py:449-453). Here the file's own little-endian sample bytes are uploaded as they are and one
 * kernel converts (libswresample's rules: u8 (x-128)/128, s16 x/2^15, s24 x/2^23, s32 x/2^31, f64 → f32) and
 * transposes them into the resident planar clip:
 *   raw_dev     [n_frames][channels] samples of `format`, interleaved, as in a WAV 'data' chunk / decoded FLAC
 *   planar_dev  [channels][clip_samples] float32; frames land at [offset, offset + n_frames) of every channel */
enum { SFB_PCM_U8 = 0, SFB_PCM_S16 = 1, SFB_PCM_S24 = 2, SFB_PCM_S32 = 3, SFB_PCM_F32 = 4, SFB_PCM_F64 = 5 };
int sfb_pcm_ingest(sfb_ctx* ctx, const void* raw_dev, int64_t n_frames, int channels, int format,
                   float* planar_dev, int64_t clip_samples, int64_t offset);

/* DynamicNumber.next (dynamics.py:197-250) over frames for any float32 vector, in place on
 * values_inout_dev [n_frames][lanes]: row k holds the target of frame k on entry and the state after frame
 * k's update on return. Same kernel as the spectrogram columns of sfb_audio_track. */
int sfb_dynamics_scan(sfb_ctx* ctx, float* values_inout_dev, int lanes, const double* dt_dev, int n_frames,
                      const sfb_dynamics_params* params);

/* ------------------------------------------------------------------------------------------------ */
/* Piano roll — replaces the per-frame Python tree walk of ShaderPiano.update (piano/module.py:202-277).
 * notes_dev: every note of the scene, sorted by pitch (stable in insertion order); pitch_offset_dev [129]:
 * notes of pitch p are notes_dev[pitch_offset[p] .. pitch_offset[p+1]). `order` is the insertion index
 * (add_note call order), which together with the integer-second bucket decides the reference's iteration
 * order, hence the roll slots and which note wins a key. */
typedef struct sfb_piano_note {
    double start, end;          /* seconds, float64 like PianoNote.start/.end */
    int32_t note, channel, velocity, order;
} sfb_piano_note;

/* All frames at once: for frame k at scene time time_dev[k] writes the key-press targets and the channel row
 * update() would build (module.py:213-247; channel -1 = not playing) and the lowest / highest pitch with
 * upcoming notes (module.py:262-266; (128, -1) when there is none). [n_frames][128], [n_frames][128], [n_frames][2] */
int sfb_piano_track(sfb_ctx* ctx, const sfb_piano_note* notes_dev, const int32_t* pitch_offset_dev,
                    const double* time_dev, int n_frames, double time_offset, double roll_time,
                    double lookup_time, double release_before_end, int global_min, int global_max,
                    float* key_target_out_dev, float* channel_out_dev, int32_t* upcoming_out_dev);
/* One frame's roll texture (module.py:222-231) into roll_out_dev [128][256][4] float32 — e.g. the storage of
 * the iPianoRoll texture: texel (slot, pitch) = (start, end, channel, velocity). *overflow_dev is set to 1 when a
 * pitch has more than 2048 visited notes in the window (slots would then be wrong). */
int sfb_piano_roll(sfb_ctx* ctx, const sfb_piano_note* notes_dev, const int32_t* pitch_offset_dev,
                   double time, double roll_time, double lookup_time, int global_min, int global_max,
                   float* roll_out_dev, int32_t* overflow_dev);

/* ------------------------------------------------------------------------------------------------ */
/* Textures — replaces moderngl Context.texture / Texture.write / .filter / .repeat_x/y as used by
 * ShaderTexture.make/apply/write (texture.py:250-283,313-325). Backed by a cudaArray + texture object
 * (hardware filtering) and a pitch-linear mirror (exact float32 filtering, TMA staging).
 * components 1..4 (3 is stored padded to 4, alpha reads 1 like GL); rows bottom-first. */
enum { SFB_DTYPE_U8 = 0, SFB_DTYPE_F32 = 1, SFB_DTYPE_F16 = 2 };
enum { SFB_FILTER_NEAREST = 0, SFB_FILTER_LINEAR = 1 };

int sfb_tex_create(sfb_ctx* ctx, int width, int height, int components, int dtype,
                   int filter, int repeat_x, int repeat_y, sfb_tex** out);
int sfb_tex_destroy(sfb_tex* tex);
int sfb_tex_set_sampling(sfb_tex* tex, int filter, int repeat_x, int repeat_y);
/* data: tightly packed w*h*components elements of the texture dtype; on_device selects the pointer kind */
int sfb_tex_write(sfb_tex* tex, const void* data, int on_device, int x, int y, int w, int h);
/* Zero-copy: sample straight from a device buffer laid out [height][width][components_padded] of the
 * texture dtype (e.g. a row of sfb_stft_mel's output) — the GL upload of texture.py:321 disappears.
 * Pass NULL to go back to the texture's own storage. */
int sfb_tex_bind_external(sfb_tex* tex, const void* data_dev);
int sfb_tex_read(sfb_tex* tex, void* data_host);          /* blocks; debugging / tests */
/* Render-target use (the FBO of texture.py:266-267): device pointer of the texture's own linear storage,
 * [height][width][padded components]. Kernels that write it make the cudaArray copy stale, so the texture
 * samples through the exact (linear-mirror) path until the next sfb_tex_write. */
int sfb_tex_storage(sfb_tex* tex, void** data_dev, size_t* bytes);
/* texture(sampler, uv) probe: n normalised coordinates uv_dev [n][2] → out_dev [n][4] float32, with the
 * texture's filter / wrap state and SFB_FILTER_EXACT or SFB_FILTER_HARDWARE (parity tests of the sampler) */
int sfb_tex_sample(sfb_tex* tex, const float* uv_dev, int n, int flags, float* out_dev);

/* ------------------------------------------------------------------------------------------------ */
/* Programs — replaces ShaderProgram.compile/use_pipeline/render (shader.py:313-405): the GLSL of the
 * in-scope scenes is transliterated ahead of time to CUDA device functions; "compile" is a lookup. */
enum {
    SFB_SCENE_DEFAULT = 0,     /* resources/shaders/fragment/default.glsl      (Basic)      */
    SFB_SCENE_SHADERTOY = 1,   /* examples/basic/shaders/shadertoy.frag        (ShaderToy)  */
    SFB_SCENE_VISUALIZER = 2,  /* examples/basic/shaders/visualizer.frag       (Visualizer) */
    SFB_SCENE_BARS = 3,        /* examples/basic/shaders/bars.frag             (MusicBars)  */
    SFB_SCENE_WAVEFORM = 4,    /* examples/basic/shaders/waveform.frag         (Waveform)   */
    SFB_SCENE_MANDELBROT = 5,  /* examples/fractals/shaders/mandelbrot.frag    (Mandelbrot) */
    SFB_SCENE_TETRATION = 6,   /* examples/fractals/shaders/tetration.frag     (Tetration)  */
    SFB_SCENE_RAYMARCH = 7,    /* examples/basic/shaders/raymarch.frag         (RayMarch)   */
    /* multi-program / multi-layer / temporal scenes */
    SFB_SCENE_MULTISHADER_CHILD = 8,  /* examples/basic/demo.py:74-79, inline      (MultiShader.child)  */
    SFB_SCENE_MULTISHADER = 9,        /* examples/basic/demo.py:83-89, inline      (MultiShader)        */
    SFB_SCENE_MULTIPASS = 10,         /* examples/basic/shaders/multipass.frag     (Multipass, 2 layers) */
    SFB_SCENE_MOTIONBLUR = 11,        /* examples/basic/shaders/motionblur.frag    (MotionBlur, temporal) */
    SFB_SCENE_DYNAMICS = 12,          /* examples/basic/demo.py:120-125, inline    (Dynamics)           */
    SFB_SCENE_AUDIO = 13,             /* examples/basic/demo.py:150-154, inline    (Audio)              */
    SFB_SCENE_LIFE_SIMULATION = 14,   /* examples/basic/shaders/life/simulation.glsl (Life.simulation)  */
    SFB_SCENE_LIFE_VISUALS = 15,      /* examples/basic/shaders/life/visuals.glsl  (Life)               */
    SFB_SCENE_PIANO = 16,             /* examples/shaders/piano.frag of THIS repository (the reference ships  */
                                      /* ShaderPiano but no fragment for it)                                 */
    SFB_SCENE_COUNT = 17,
};

#define SFB_MAX_EXTRA 16
#define SFB_MAX_SAMPLERS 20

/* One POD block passed by value to the kernel (__grid_constant__): the uniforms of
 * ShaderScene.pipeline (scene.py:687-703), ShaderCamera.pipeline + its ShaderDynamics (camera.py:146-201),
 * then `extra` slots for module / user uniforms in the order sfb_scene_info reports. Values are the
 * float32 casts GL would make at upload. */
typedef struct sfb_uniforms {
    float iTime, iTau, iDuration, iDeltatime;
    float iResolution[2];
    float iWantAspect, iQuality, iSSAA, iFramerate;
    int32_t iFrame, iRealtime, iLayer, iMouseInside;
    float iMouse[2];
    int32_t iMouse1, iMouse2;
    int32_t iCameraMode, iCameraProjection;
    float iCameraPosition[3], iCameraRight[3], iCameraUpward[3], iCameraForward[3], iCameraZenith[3];
    float iCameraZoom, iCameraIsometric, iCameraFocalLength, iCameraOrbital, iCameraDolly, iCameraSeparation;
    float extra[SFB_MAX_EXTRA][4];
} sfb_uniforms;

typedef struct sfb_scene_info {
    const char* name;                       /* "visualizer", ... */
    const char* reference;                  /* reference file the kernel transliterates */
    int n_extra;    const char* extra[SFB_MAX_EXTRA];       /* uniform names → extra[i] slots */
    int n_samplers; const char* samplers[SFB_MAX_SAMPLERS]; /* sampler2D names → samplers[i] */
    int n_required;                         /* the first n_required samplers must be bound; the rest are the  */
                                            /* history of a temporal texture, as deep as the scene made it     */
} sfb_scene_info;

int sfb_scene_lookup(const char* name, int* scene);       /* SFB_ENOTFOUND when not built in */
int sfb_scene_info_get(int scene, sfb_scene_info* info);

/* Render flags */
enum {
    SFB_FILTER_EXACT = 0,      /* float32 bilinear on point-fetched texels (parity default)           */
    SFB_FILTER_HARDWARE = 1,   /* cudaTextureObject filtering (9-bit weights, like GL hardware)       */
    SFB_RENDER_LITERAL = 2,    /* force the literal transliteration of the GLSL (no scene-specific fast  */
                               /* path); the parity anchor the optimised kernels are compared against   */
    SFB_RENDER_TILED = 4,      /* visualizer: skip the separable kernel, use the per-pixel tiled one      */
};

/* iScreen pass (K3): one thread per fragment of a target_w x target_h RGBA8 target — replaces
 * ShaderProgram.render_to_fbo's fullscreen TRIANGLE_STRIP draw (shader.py:367-375,398-403).
 * dst_rgba8_dev: target_w*target_h*4 bytes. dst_f32_dev: optional float4 per fragment, the colour
 * BEFORE the 8-bit store (parity tests); NULL normally. */
int sfb_render_screen(sfb_ctx* ctx, int scene, const sfb_uniforms* uniforms,
                      sfb_tex* const* samplers, int n_samplers, int flags,
                      int target_w, int target_h, void* dst_rgba8_dev, float* dst_f32_dev);

/* The same pass into a texture's own storage, in the texture's format (u8 or f32, 1-4 components): the FBO
 * render of a child program / a layer / a temporal slot (shader.py:398-405, texture.py:266-267). The target
 * then samples through the exact path until the next sfb_tex_write. iResolution comes from `uniforms`. */
int sfb_render_target(sfb_ctx* ctx, int scene, const sfb_uniforms* uniforms,
                      sfb_tex* const* samplers, int n_samplers, int flags, sfb_tex* target);

/* Final pass (K4): fragment/final.glsl:3-33 over an RGBA8 iScreen (LINEAR, CLAMP_TO_EDGE) into a
 * width x height target with `components` 3 (rgb24, what ffmpeg gets: exporting.py:94-103) or 4. */
int sfb_render_final(sfb_ctx* ctx, const void* screen_rgba8_dev, int screen_w, int screen_h,
                     int width, int height, int subsample, int components, void* dst_dev);

/* Fused K3+K4: shades the ssaa x ssaa sub-samples of each output pixel, quantises each to 8 bit
 * exactly like the RGBA8 iScreen store, box-averages and quantises again. Bit-identical to
 * screen+final whenever final.glsl degenerates to a box filter: integer ssaa with
 * subsample == ssaa or 2*subsample == ssaa; other combinations → SFB_EINVAL.
 * The iScreen texture never touches HBM. */
int sfb_render_frame(sfb_ctx* ctx, int scene, const sfb_uniforms* uniforms,
                     sfb_tex* const* samplers, int n_samplers, int flags,
                     int width, int height, int ssaa, int subsample, int components, void* dst_dev);

/* Parity probe of the fused path: the same kernels sfb_render_frame would pick (same flags, same launch) also
 * store every shaded sub-sample's fragColor BEFORE the 8-bit iScreen store as float4 into screen_f32_dev
 * [height*ssaa][width*ssaa][4] — what shader.py:367-375 would leave in a float FBO. Tests hold the production
 * kernels to the reference GLSL's values at 1e-3 per channel with it; exports never call it. */
int sfb_render_frame_probe(sfb_ctx* ctx, int scene, const sfb_uniforms* uniforms,
                           sfb_tex* const* samplers, int n_samplers, int flags,
                           int width, int height, int ssaa, int subsample, int components, void* dst_dev,
                           float* screen_f32_dev);

/* Which kernel sfb_render_frame picks for the visualizer scene (no GPU needed; diagnostics and tests):
 * *rows_per_thread = 8 or 4 when the separable kernel (one thread per fragment column) shades the frame and
 * *window_rows = background rows it stages per CTA; 0 when the per-pixel tiled kernel does (rotated / stereo /
 * equirectangular camera, ssaa 3, or a vertical texel step per fragment too coarse for 4 texel rows). */
int sfb_visualizer_plan(const sfb_uniforms* uniforms, int background_w, int background_h,
                        int width, int height, int ssaa, int* rows_per_thread, int* window_rows);

/* ------------------------------------------------------------------------------------------------ */
/* FLAC streams — audio files without an ffmpeg child (the reference decodes every compressed file through
 * `ffmpeg -f f32le`, ffmpeg.py:1240-1333). Host-side: the bit stream is serial work, no GPU and no sfb_ctx involved.
 *   sfb_flac_info_get  STREAMINFO of a stream held in memory (an ID3v2 tag in front is skipped).
 *   sfb_flac_decode    every frame → interleaved samples [frame][channel], right-justified int32 (a 16-bit stream
 *                      yields values in [-32768, 32767]); pcm NULL = count only. All CRC-8 / CRC-16 are verified
 *                      (SFB_EINVAL on a mismatch or malformed stream); md5 is for the caller to check. */
typedef struct sfb_flac_info {
    int32_t samplerate, channels, bits_per_sample, has_md5;
    int64_t total_samples;                  /* per channel; 0 = not recorded in the stream */
    int32_t min_block, max_block;
    uint8_t md5[16];                        /* of the interleaved little-endian samples, ceil(bits/8) bytes each */
} sfb_flac_info;
int sfb_flac_info_get(const void* data, size_t bytes, sfb_flac_info* info);
int sfb_flac_decode(const void* data, size_t bytes, int32_t* pcm, int64_t capacity_frames, int64_t* decoded_frames);

/* ------------------------------------------------------------------------------------------------ */
/* Video frames as textures — replaces ShaderVideo.update's host-side work (video.py:57-66: np.flip, a C-order copy,
 * texture.write of an rgb24 frame ffmpeg decoded). frame_dev holds ONE frame in the file's own layout (rows top to
 * bottom when top_down, as decoders and Y4M deliver them); the kernel flips it to GL's bottom-row-first order,
 * converts planar YUV to RGB (BT.601 limited range; SFB_VIDEO_FULL_RANGE = JPEG range) and writes the texture's
 * storage (8-bit, 3 or 4 components, width x height). sfb_video_frame_bytes = size of one frame (0: unknown format). */
enum {
    SFB_VIDEO_RGB24 = 0, SFB_VIDEO_RGBA32 = 1,
    SFB_VIDEO_YUV420P = 2, SFB_VIDEO_YUV422P = 3, SFB_VIDEO_YUV444P = 4,   /* planar Y, U, V (I420 order) */
    SFB_VIDEO_FULL_RANGE = 0x100,                                          /* or'ed to a YUV format */
};
size_t sfb_video_frame_bytes(int format, int width, int height);
int sfb_video_frame(sfb_ctx* ctx, const void* frame_dev, int format, int width, int height, int top_down, sfb_tex* texture);

/* ------------------------------------------------------------------------------------------------ */
/* Run-time compiled programs — the other half of ShaderProgram.compile (shader.py:313-349 hands ANY fragment text to
 * the GL driver). A fragment this library has no ahead-of-time kernel for is translated GLSL → CUDA by the host side
 * (shaderflow_b200/glsl) and compiled here:
 *   sfb_jit_compile   NVRTC → a SASS image for sm_100a. `source` includes its headers by name from the in-memory table
 *                     (header_names[i] → header_sources[i]). Needs no GPU; libnvrtc.so.12 is loaded on first use
 *                     (SFB_ENOTFOUND when absent). *log (optional) receives the compiler's diagnostics, also on
 *                     success; release image and log with sfb_jit_free. SFB_EINVAL = the source does not compile.
 *   sfb_program_load  loads the image on ctx's device → *scene >= SFB_SCENE_PROGRAM_BASE, accepted by sfb_render_screen,
 *                     sfb_render_target, sfb_render_frame and sfb_render_frame_probe like a built-in scene (the image
 *                     must define the kernels sfb_jit_screen and sfb_jit_frame, csrc/jit/jit_kernels.cuh);
 *                     n_samplers = samplers the program reads, in the order the caller will bind them. */
#define SFB_SCENE_PROGRAM_BASE 1000
enum { SFB_JIT_FMAD = 1 };   /* allow a*b+c → fma contraction (default: one rounding per operation, like the checker) */
int sfb_jit_compile(const char* source, const char* const* header_names, const char* const* header_sources, int n_headers,
                    int flags, void** image, size_t* image_bytes, char** log);
void sfb_jit_free(void* image_or_log);
int sfb_program_load(sfb_ctx* ctx, const void* image, size_t image_bytes, int n_samplers, int* scene);
int sfb_program_unload(sfb_ctx* ctx, int scene);

/* ------------------------------------------------------------------------------------------------ */
/* Frame sink — replaces turbopipe.pipe/sync/done + fbo.read_into (exporting.py:140-174): a ring of
 * n_buffers device frames, each paired with a pinned host buffer; submit enqueues an async D2H on a copy
 * stream ordered after the render stream, a writer thread does write(fd) in submission order.
 * fd < 0 is a null sink (frames are copied to the host and dropped).
 *   acquire: device pointer of the next ring frame to render into; blocks while its previous contents
 *            are still being copied / written (the back-pressure of exporting.py:168's turbopipe.sync).
 *   submit:  frame_dev NULL or the acquired pointer → zero-copy; any other device pointer is first
 *            copied D2D into the ring frame on the render stream. */
int sfb_pipe_open(sfb_ctx* ctx, int fd, int n_buffers, size_t frame_bytes, sfb_pipe** out);
int sfb_pipe_acquire(sfb_pipe* pipe, void** frame_dev);
int sfb_pipe_submit(sfb_pipe* pipe, const void* frame_dev);
/* Re-targets an idle ring (everything submitted has been written) at another file descriptor, so one ring —
 * its pinned buffers, copy stream and writer thread — serves consecutive exports. */
int sfb_pipe_set_fd(sfb_pipe* pipe, int fd);
int sfb_pipe_sync(sfb_pipe* pipe);                       /* blocks until every submitted frame is written */
int sfb_pipe_stats(sfb_pipe* pipe, uint64_t* frames, uint64_t* bytes);
int sfb_pipe_close(sfb_pipe* pipe);

/* ------------------------------------------------------------------------------------------------ */
/* Sink of a frame-SHARDED export (same role as the ring above, exporting.py:140-174, when the
 * export runs as one process per GPU). One shared-memory segment holds a host ring of `slots` frames per rank;
 * every rank pins its own ring and copies its frames device → host over ITS OWN PCIe link while it shades;
 * ownership is block-cyclic (frame g belongs to rank (g / block) % world); rank 0's writer thread walks the
 * frames in time order and write()s them to the sink's fd. Cross-process state = two counters per rank in the
 * segment header; no collective, no kernel.
 *   open:   rank 0 passes segment_path NULL and creates the segment; the others attach to sfb_sink_path(rank 0's).
 *           ctx NULL = host-only sink (frames handed over as host memory; CPU tests of the protocol).
 *   begin:  every rank, then the caller synchronises the ranks before the first frame (rank 0 resets the
 *           counters and starts the writer; fd < 0 = null sink).
 *   acquire / submit: like sfb_pipe_*, for the frames THIS rank owns, in time order.
 *   finish: blocks until this rank's frames have left (rank 0: until the whole export is written).
 *   abort:  unblocks every rank with an error. */
typedef struct sfb_sink sfb_sink;
int sfb_sink_open(sfb_ctx* ctx, const char* segment_path, int rank, int world, int slots,
                  size_t frame_bytes, sfb_sink** out);
const char* sfb_sink_path(sfb_sink* sink);
int sfb_sink_begin(sfb_sink* sink, int64_t total_frames, int block, int fd);
int sfb_sink_owner(sfb_sink* sink, int64_t frame, int* owner);
int sfb_sink_acquire(sfb_sink* sink, void** frame_dev);
int sfb_sink_submit(sfb_sink* sink);
int sfb_sink_submit_host(sfb_sink* sink, const void* frame_host);
int sfb_sink_finish(sfb_sink* sink, uint64_t* frames, uint64_t* bytes);
int sfb_sink_abort(sfb_sink* sink);
int sfb_sink_close(sfb_sink* sink);

#ifdef __cplusplus
}
#endif
#endif /* SFB200_H */
