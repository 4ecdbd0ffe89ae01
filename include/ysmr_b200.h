/* ysmr_b200 -- C-ABI of the B200 (sm_100a) detect + link hot path of YSMR.
 *
 * The reference (schwanbeck/YSMR) is pure Python and has no FFI of its own; the seam this library drops into is
 * the body of the per-frame loop of track_bacteria() (ysmr/track_eval.py:156-316).  Each entry point below names
 * the reference lines it replaces.  Conventions:
 *   - plain C, no exceptions across the boundary; every function returns 0 (YSMR_OK) or a negative YSMR_E_* code,
 *     a human-readable message is available from ysmr_last_error();
 *   - pointers named d_* are DEVICE pointers owned by the caller (e.g. torch tensors), h_* are HOST pointers;
 *     the library never frees caller memory;
 *   - one context per GPU; a context is not thread-safe, different contexts may be driven from different threads;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are asynchronous on it
 *     unless stated otherwise.
 * The binding a YSMR maintainer would add (ctypes) is shown in INTEGRATION.md and implemented in ysmr_b200/_lib.py.
 */
#ifndef YSMR_B200_H
#define YSMR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YSMR_ABI_VERSION 5

enum {
    YSMR_OK = 0,
    YSMR_E_INVALID = -1,      /* bad argument / unsupported geometry */
    YSMR_E_CUDA = -2,         /* CUDA runtime error, text in ysmr_last_error */
    YSMR_E_OVERFLOW = -3,     /* a per-frame capacity (runs, blobs, tracks, contour points, rows) was exceeded */
    YSMR_E_STATE = -4,        /* call sequence error (e.g. link before gains were set) */
    YSMR_E_NOMEM = -5
};

/* Bits of the device-side status word returned by ysmr_status(). */
enum {
    YSMR_ST_RUN_OVERFLOW = 1,      /* more foreground runs in a frame than max_runs */
    YSMR_ST_BLOB_OVERFLOW = 2,     /* more external blobs in a frame than max_blobs */
    YSMR_ST_POINT_OVERFLOW = 4,    /* contour longer than the contour scratch pool allows */
    YSMR_ST_TRACK_OVERFLOW = 8,    /* more live tracks than max_tracks */
    YSMR_ST_ROW_OVERFLOW = 16,     /* rows_capacity of ysmr_link exceeded */
    YSMR_ST_GATE_TIMEOUT = 32      /* ysmr_track_*: a linker kernel waited ~10 s of SM clocks for its chunk's detections */
};

/* Threshold modes of the loop (track_eval.py:185, 198, 219). */
enum {
    YSMR_MODE_ADAPTIVE_DOUBLE = 0, /* 'adaptive double threshold' > 0 : mask + markers + binary_propagation */
    YSMR_MODE_ADAPTIVE_SINGLE = 1, /* == 0 : one adaptive threshold                                      */
    YSMR_MODE_MEAN_STD = 2         /* <  0 : moving average of (mean +- std +- offset), cv2.threshold      */
};

/* The tracking.ini keys the hot path reads, plus capacities. */
typedef struct ysmr_params {
    int32_t white_on_dark;   /* 'white bacteria on dark background' (track_eval.py:127)                   */
    int32_t offset;          /* 'threshold offset for detection' as written in the ini (NOT pre-negated)  */
    double  adt;             /* 'adaptive double threshold' (track_eval.py:185,198,206)                   */
    double  fps;             /* fps_of_file: CentroidTracker(max_disappeared=fps, fps=fps) (track_eval.py:109-116) */
    int32_t use_gsff;        /* not 'disable gsff'                                                        */
    int32_t n_f;             /* 'number of LSFFs' (<= YSMR_MAX_FILTERS)                                   */
    int32_t n_min;           /* 'minimum horizon size'                                                    */
    int32_t n_max;           /* 'maximum horizon size' (<= YSMR_MAX_HORIZON)                              */
    int32_t max_blobs;       /* capacity: detections per frame                                            */
    int32_t max_tracks;      /* capacity: simultaneously live tracks                                      */
    int32_t max_runs;        /* capacity: foreground runs per frame in the labelling kernel               */
    int32_t max_batch;       /* capacity: frames per ysmr_detect call                                     */
    double  max_distance;    /* association gate; <= 0 means off.  The reference has NO gate
                                (tracker.py:171-177), so anything but off diverges from it.              */
    int32_t reserved[4];     /* reserved[0] > 0: explicit GSFF horizons n_i[0..n_f-1] (gsff.py:103-106 evaluated by the host
                                with a FLOAT n_max when 'maximum horizon size' is unset, tracker.py:58-59); else n_min/n_max */
} ysmr_params;

#define YSMR_MAX_FILTERS 4
#define YSMR_MAX_HORIZON 64

/* One output row = one line of <video>_list.csv before sorting (track_eval.py:313-316, helper_file.py:1456-1477):
 * TRACK_ID, POSITION_T, POSITION_X, POSITION_Y, WIDTH, HEIGHT, DEGREES_ANGLE. */
typedef struct ysmr_row {
    int32_t frame;           /* POSITION_T  (curr_frame_count)                  */
    int32_t track_id;        /* TRACK_ID                                        */
    double  x, y;            /* GSFF-corrected position (tracker.py:222)        */
    float   w, h, deg;       /* cv2.minAreaRect size/angle, or 0,0,0 when the track was not matched */
    int32_t pad;
} ysmr_row;                  /* 40 bytes */

/* Optional per-stage outputs of ysmr_detect for parity tests; any pointer may be NULL.
 * All are DEVICE pointers to n_frames x H x W bytes ({0,255} for the masks), except first_xy. */
typedef struct ysmr_debug_out {
    uint8_t *d_grey;         /* cv2.cvtColor           (track_eval.py:180)      */
    uint8_t *d_blurred;      /* cv2.GaussianBlur 3x3   (track_eval.py:182)      */
    uint8_t *d_mean;         /* adaptiveThreshold's rounded Gaussian mean       */
    uint8_t *d_mask;         /* first adaptiveThreshold (track_eval.py:189-197) */
    uint8_t *d_markers;      /* second adaptiveThreshold (track_eval.py:200-208)*/
    uint8_t *d_out;          /* image handed to findContours (track_eval.py:211-214 / 248-253) */
    int32_t *d_first_xy;     /* n_frames x max_blobs x 2: raster-first pixel of each contour, cv2 order */
    int32_t *d_scalar_thr;   /* n_frames: curr_threshold of the mean/std mode (track_eval.py:236) */
} ysmr_debug_out;

typedef struct ysmr_ctx ysmr_ctx;

/* Library-wide. */
int         ysmr_abi_version(void);
const char *ysmr_last_error(const ysmr_ctx *ctx);      /* ctx may be NULL: last error of a failed ysmr_create */

/* Fill `p` with the reference defaults (helper_file.py:160-282) and default capacities. */
void ysmr_default_params(ysmr_params *p);

/* Context for frames of height x width x channels (channels: 3 = BGR as cap.read() delivers, track_eval.py:159;
 * 1 = grey plane).  Replaces the per-video setup of track_eval.py:109-132. */
int ysmr_create(ysmr_ctx **out, int device, int height, int width, int channels, const ysmr_params *params);
int ysmr_destroy(ysmr_ctx *ctx);

/* Gain matrix of least-squares FIR filter `filter` (0-based) with horizon `horizon`: 4 x (2*horizon) doubles,
 * row-major, HOST memory, computed like gsff.py:112-153 (the Python host does it with the same numpy calls as the
 * reference so the values are bit-identical).  Must be called for every filter before ysmr_link when use_gsff. */
int ysmr_set_gsff_gain(ysmr_ctx *ctx, int filter, int horizon, const double *h_gain);

/* Detection of n_frames frames resident in HBM (replaces track_eval.py:180-303 for each of them).
 *   d_frames      n_frames frames of H*W*channels bytes, frame i at d_frames + i*frame_stride
 *   d_blob_count  [n_frames] number of min-area rectangles of each frame
 *   d_blobs       [n_frames][max_blobs][5] float: cx, cy, w, h, deg -- cv2.minAreaRect values in cv2.findContours
 *                 order (= the `rects` list handed to ct.update, track_eval.py:285-311)
 *   dbg           optional per-stage outputs (NULL in production)
 * first_frame is the index of frames[0] in the video (only the mean/std mode depends on it). */
int ysmr_detect(ysmr_ctx *ctx, const uint8_t *d_frames, int n_frames, int64_t frame_stride, int first_frame,
                int32_t *d_blob_count, float *d_blobs, const ysmr_debug_out *dbg, void *stream);

/* Sequential linking of n_frames frames (replaces ct.update + the row append, track_eval.py:311-316,
 * tracker.py:93-230, gsff.py:204-347).  Stateful and resumable: call it chunk after chunk in frame order.
 *   d_rows / rows_capacity   output rows in emission order (frame-major, tracks in insertion order)
 *   d_n_rows                 [1] number of rows written by THIS call */
int ysmr_link(ysmr_ctx *ctx, const int32_t *d_blob_count, const float *d_blobs, int first_frame, int n_frames,
              ysmr_row *d_rows, int64_t rows_capacity, int64_t *d_n_rows, void *stream);
/* Same, appending: rows are written from d_rows[*d_n_rows] on and *d_n_rows (device, in/out) becomes the running total, so
 * chunk after chunk can be enqueued without a host round trip (streamed multi-GPU hand-over, bench.py). */
int ysmr_link_append(ysmr_ctx *ctx, const int32_t *d_blob_count, const float *d_blobs, int first_frame, int n_frames,
                     ysmr_row *d_rows, int64_t rows_capacity, int64_t *d_n_rows, void *stream);
int ysmr_link_reset(ysmr_ctx *ctx);                    /* forget all tracks, next id = 0 (new video) */

/* Snapshot of the linker (for tests and for hand-over between processes).  Synchronous.
 * Pass h_buf == NULL to query the size. */
int ysmr_link_state_export(ysmr_ctx *ctx, void *h_buf, size_t *size);
int ysmr_link_state_import(ysmr_ctx *ctx, const void *h_buf, size_t size);
int ysmr_link_live_tracks(ysmr_ctx *ctx, int32_t *n_live, int32_t *next_id);   /* synchronous */

/* Synchronise `stream` and return the device-side status bits (YSMR_ST_*), clearing them.  *first_bad_frame
 * (may be NULL) receives the lowest frame index that raised a bit, or -1. */
int ysmr_status(ysmr_ctx *ctx, void *stream, int32_t *status_bits, int32_t *first_bad_frame);

/* End-to-end convenience over HOST buffers: n_frames frames in (preferably pinned) host memory are copied to the
 * GPU in chunks on a copy stream, detected and linked while the next chunk is in flight, and the rows are copied
 * back.  Synchronous.  This is the call the track_bacteria() drop-in makes per decoded chunk. */
int ysmr_track_host(ysmr_ctx *ctx, const uint8_t *h_frames, int n_frames, int64_t frame_stride, int first_frame,
                    ysmr_row *h_rows, int64_t rows_capacity, int64_t *n_rows);

/* Same pipeline over frames already resident in HBM; rows stay on the device.  Asynchronous on `stream`
 * (internally forks a second stream so the linker of chunk i overlaps detection of chunk i+1). */
int ysmr_track_device(ysmr_ctx *ctx, const uint8_t *d_frames, int n_frames, int64_t frame_stride, int first_frame,
                      ysmr_row *d_rows, int64_t rows_capacity, int64_t *d_n_rows, void *stream);

/* Row sink on the device (replaces helper_file.sort_list, /root/reference/ysmr/helper_file.py:1538-1574, for the rows
 * of the hot loop): with ysmr_rows_archive(ctx, 1) every ysmr_track_host call also appends its rows to an archive in
 * device memory (h_rows may then be NULL: nothing is copied back per call).  ysmr_rows_sorted groups the archive by
 * (track_id, frame) -- the order of the final <video>_list.csv -- with a counting sort on the GPU and copies it to the
 * host ONCE; with h_rows == NULL it only reports the count.  ysmr_rows_archive(ctx, 0) / ysmr_link_reset drop it. */
int ysmr_rows_archive(ysmr_ctx *ctx, int enabled);
int ysmr_rows_sorted(ysmr_ctx *ctx, ysmr_row *h_rows, int64_t rows_capacity, int64_t *n_rows);

/* ---- Track selection: the data-parallel body of select_tracks() / find_good_tracks(),
 * /root/reference/ysmr/track_eval.py:408-843, on the rows of <video>_list.csv.  Values are the reference's settings AFTER
 * get_configs() (helper_file.py:777-786: percentages already divided by 100, 'maximal empty frames in %' / 100 + 1). */
typedef struct ysmr_select_params {
    double  area_lo, area_hi;       /* 'extreme area outliers lower / upper end in px*px' (track_eval.py:624-630)       */
    double  area_factor;            /* 'exclude measurement when above x times average area', 0 = off (:632-637)         */
    double  q_area;                 /* 'percent quantiles excluded area' (fraction), <= 0 = off (:703-712)               */
    double  stop_outliers_above;    /* 'stop excluding motility outliers if total count above percent' (fraction, :728) */
    double  max_empty;              /* 'maximal empty frames in %' (:471)                                                 */
    double  ratio_min, ratio_max;   /* 'average width/height ratio min. / max.' (:478-480)                                */
    double  edge;                   /* 'percent of screen edges to exclude' (fraction, :483-494)                          */
    int32_t min_len_frames;         /* int(round(fps) * 'minimal length in seconds') (:586)                               */
    int32_t limit_frames;           /* int(round(fps) * 'limit track length to x seconds'), 0 = off (:587, 781)           */
    int32_t limit_exactly;          /* 'limit track length exactly' (:785-790)                                            */
    int32_t omit_motility_outliers; /* 'try to omit motility outliers' (:713)                                             */
    int32_t max_holes;              /* 'maximal consecutive holes' (:462)                                                 */
    int32_t max_recursion;          /* 'maximal recursion depth' (:507-509), 0 .. 4000                                    */
    int32_t frame_h, frame_w;
} ysmr_select_params;

enum {                              /* indices of the info array of ysmr_select_tracks */
    YSMR_SI_STATUS = 0,             /* YSMR_SEL_* */
    YSMR_SI_ROWS_BEFORE = 1, YSMR_SI_ROWS_AFTER = 2, YSMR_SI_TRACKS_BEFORE = 3, YSMR_SI_TRACKS_AFTER = 4,   /* log line :686-691 */
    YSMR_SI_Q1_AREA = 5, YSMR_SI_Q3_AREA = 6, YSMR_SI_Q1_DIST = 7, YSMR_SI_Q3_DIST = 8, YSMR_SI_FENCE = 9,
    YSMR_SI_OUTLIERS = 10, YSMR_SI_OUTLIERS_OFF = 11, YSMR_SI_GOOD_TRACKS = 12, YSMR_SI_LAUNCHES = 13,
    YSMR_SELECT_INFO = 16
};
enum {
    YSMR_SEL_OK = 0,
    YSMR_SEL_TOO_SHORT_BEFORE = 1,  /* fewer rows than the minimal length before the clean-up: the reference returns None (:599-606) */
    YSMR_SEL_TOO_SHORT_AFTER = 2,   /* ... after the clean-up (:676-684) */
    YSMR_SEL_NO_TRACKS = 3          /* no track passed (:816-819) */
};

/* n_rows rows grouped by (TRACK_ID, POSITION_T) in HOST memory, columns as the reference's data frame holds them
 * (uint32, uint32, float64 x 4).  Outputs (HOST): h_good[n_rows] = 1 for the rows of the selected fragments
 * (df['good_track'], :822-828); h_clean_index[n_rows] = the row's index after the initial clean-up (the 'index' column
 * reset_index leaves in the returned frame, :834) or -1 if the clean-up dropped it; kick_reasons[9] (:749, 796-812);
 * info[YSMR_SELECT_INFO].  Synchronous; runs on `device`. */
int ysmr_select_tracks(int device, int64_t n_rows, const uint32_t *h_track_id, const uint32_t *h_t, const double *h_x,
                       const double *h_y, const double *h_w, const double *h_h, const ysmr_select_params *params,
                       uint8_t *h_good, int32_t *h_clean_index, int64_t *kick_reasons, double *info);
const char *ysmr_select_last_error(void);

/* ---- Per-track statistics (PARTIAL): the columns of evaluate_tracks()' df_stats that are reductions
 * over a track's rows, /root/reference/ysmr/track_eval.py:905-945, 1030-1096.  Input: the selected rows (what select_tracks
 * returns), grouped by track, HOST columns; h_track_start[n_tracks] = first row of every track.  median_kernel = the second
 * median-filter size of `moving` (:933-940: round(fps), made odd).  Output h_stats[n_tracks][YSMR_STAT_COLUMNS] (HOST). */
enum {
    YSMR_STAT_DISTANCE = 0,       /* 'Distance (um)'  groupby sum of travelled_dist (pandas' Kahan sum)        (:1047) */
    YSMR_STAT_SPEED = 1,          /* 'Speed (um/s)'   distance / time, 0 for tracks that never move            (:1052-1055) */
    YSMR_STAT_TIME = 2,           /* 'Time (s)'       (last t_norm + 1) / fps                                   (:1046) */
    YSMR_STAT_DISPLACEMENT = 3,   /* 'Displacement (um)'  largest pairwise distance, scipy pdist().max()       (:1031) */
    YSMR_STAT_PERC_MOTILE = 4,    /* 'Perc. Motile'   sum of the twice median-filtered `moving` / frames * 100  (:1044-1045) */
    YSMR_STAT_ACR = 5,            /* 'Arc-Chord Ratio'                                                           (:1048-1062) */
    YSMR_STAT_BAC_LENGTH = 6,     /* 'Bacteria Length'  float32 mean of the float16 column bac_length          (:923, 1085) */
    YSMR_STAT_DISPL_BY_LENGTH = 7,/* 'Displacement divided by length'                                           (:1086-1092) */
    YSMR_STAT_COLUMNS = 8
};
int ysmr_track_statistics(int device, int64_t n_rows, const uint32_t *h_track_id, const uint32_t *h_t, const double *h_x,
                          const double *h_y, const double *h_w, const double *h_h, double px_per_um, double fps, int median_kernel,
                          const int32_t *h_track_start, int32_t n_tracks, double *h_stats);
const char *ysmr_statistics_last_error(void);

/* Development / measurement switches (not needed in production).  YSMR_OPT_FRONTEND_GEN: 4 (default) = the fused
 * bound-and-refine front-end kernel where it applies, 3 = always the three-kernel front-end of ABI 2 (bench.py's A/B
 * figure, and tests that hold one against the other). */
enum { YSMR_OPT_FRONTEND_GEN = 1 };
int ysmr_set_option(ysmr_ctx *ctx, int option, int value);

/* Number of kernels launched by this context since creation (bench.py's gpu_launches claim). */
int64_t ysmr_launch_count(const ysmr_ctx *ctx);

/* Per-kernel timing for bench.py's roofline: when enabled every kernel launch is bracketed by CUDA events on the stream
 * it is launched on.  ysmr_get_profile synchronises, writes the summed milliseconds and the number of launches per
 * kernel kind (arrays of YSMR_PROF_KINDS entries) and clears the record. */
enum {
    YSMR_PROF_FRONTEND = 0,   /* all front-end launches of a batch together (K1a + margins + K1b + K1c) */
    YSMR_PROF_LABEL = 1, YSMR_PROF_GEOMETRY = 2, YSMR_PROF_LINK = 3,
    YSMR_PROF_K1A = 4,        /* blur pre-pass + margins */
    YSMR_PROF_K1B = 5,        /* Gaussian + threshold decisions */
    YSMR_PROF_K1C = 6,        /* mask packing */
    YSMR_PROF_KINDS = 7
};
int ysmr_set_profiling(ysmr_ctx *ctx, int enabled);
int ysmr_get_profile(ysmr_ctx *ctx, double *ms, int64_t *launches);
/* With ysmr_set_profiling(ctx, 3) the linker's shared-memory path also accumulates SM cycles per phase (clock64 of
 * thread 0): out16[0..8] = the nine barrier-delimited phases of a frame, out16[12] = frames.  Clears the counters. */
int ysmr_link_phase_cycles(ysmr_ctx *ctx, int64_t *out16);

#ifdef __cplusplus
}
#endif
#endif /* YSMR_B200_H */
