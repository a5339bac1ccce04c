/* auvpf.h -- C ABI of the B200-native auv_particle_filter measurement cycle.
 *
 * This is the drop-in boundary for the ONE hot path of smarc_navigation that
 * this repo replaces: the particle-filter cycle of
 * auv_particle_filter/scripts/{auv_pf,auv_particle,resampling}.py.  The
 * reference has no FFI for this path (it is three Python methods on one
 * object); every entry point below names the reference
 * lines it replaces.  A maintainer binds these with ctypes from the ROS node
 * (INTEGRATION.md shows the stub); smarc_navigation_b200/filter.py is that
 * binding plus a class with the reference's predict/update/resample methods.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++/torch types.
 *  - Every function returns an int status (AUVPF_OK == 0, negative = error);
 *    auvpf_last_error() returns a thread-local message for the last failure.
 *  - Every data pointer may be a HOST pointer or a CUDA DEVICE pointer of the
 *    handle's device (detected with cudaPointerGetAttributes), so torch tensors
 *    can be passed zero-copy through data_ptr().
 *  - All work of a handle is enqueued on the handle's own CUDA stream in call
 *    order.  Calls that copy results to host memory synchronise that stream;
 *    the others return after enqueueing (auvpf_sync() waits).  A handle is not
 *    re-entrant: serialise calls per handle (the Python facade holds a lock).
 *  - Pose layout across the ABI: row-major (n, 6) float64
 *    [x, y, z, roll, pitch, yaw] in the odom frame, exactly Particle.p_pose
 *    (auv_particle.py:24,36).  Covariance arguments are the reference's
 *    6-vectors of VARIANCES in that order (auv_pf.launch:15-19).
 *  - F independent filters of M particles each live in one handle (F = 1 for
 *    the reference's single filter); particle i of filter f is row f*M + i.
 */
#ifndef AUVPF_H
#define AUVPF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AUVPF_VERSION 202 /* 201: + auvpf_pose_array_topk, auvpf_kld_bins, auvpf_kld_target, auvpf_set_next_particle_count, auvpf_particle_capacity; 202: + auvpf_shard_pose_sums, auvpf_group_pose_sums */

/* status codes */
#define AUVPF_OK 0
#define AUVPF_ERR_BADARG (-1)
#define AUVPF_ERR_CUDA (-2)
#define AUVPF_ERR_STATE (-3)      /* e.g. update_mbes before set_dem/set_beams */
#define AUVPF_ERR_DEGENERATE (-4) /* all weights zero / -inf / NaN: resample skipped */
#define AUVPF_ERR_NOMEM (-5)

/* resampler selection for auvpf_resample* */
#define AUVPF_SYSTEMATIC 0 /* resampling.py:135-168 systematic_resample            */
#define AUVPF_RESIDUAL 1   /* resampling.py:27-76 residual_resample, the one auv_pf.resample
                              calls (:182): floor(N w) copies bit-exact; the remaining draw
                              is a multinomial on the fractional parts, or -- with
                              AUVPF_RS_STRICT_REF -- the reference's literal :70-74     */

/* flags for auvpf_resample* */
#define AUVPF_RS_SORTED 0      /* new[i] = old[ancestor[i]] (ancestors ascending)    */
#define AUVPF_RS_INPLACE_REF 1 /* reference slot map, auv_pf.py:183-198: survivors keep
                                  their slot, pose[lost[r]] <- pose[dupes[r]]        */
#define AUVPF_RS_STRICT_REF 2  /* AUVPF_RESIDUAL only: reproduce resampling.py:70-74 literally
                                  (residual = w - copies, sequential sum / cumsum, NumPy's
                                  hinted searchsorted on the non-monotone result) for lock-step
                                  parity at the reference's particle counts (<= 65536)   */

#define AUVPF_RS_ASYNC 4       /* return after enqueueing: do not wait for the degenerate-weights
                                  flag (a filter without any finite weight keeps its state; query
                                  with auvpf_resample_status).  Ignored when ancestors are read
                                  back or with AUVPF_RS_INPLACE_REF.                          */

/* kernel classes for auvpf_profile_read */
enum {
    AUVPF_K_PROPAGATE = 0,
    AUVPF_K_UPDATE_MBES = 1,
    AUVPF_K_UPDATE_GPS = 2,
    AUVPF_K_NORMALISE = 3,   /* k_weights: integer weights, chunk / tile sums, total   */
    AUVPF_K_SCAN = 4,        /* scans of the residual resampler / reference slot map   */
    AUVPF_K_SELECT_GATHER = 5, /* ancestor select + particle gather (+ roughening)   */
    AUVPF_K_ROUGHEN = 6,
    AUVPF_K_ESTIMATE = 7,
    AUVPF_K_OTHER = 8,
    AUVPF_K_COUNT = 9
};

typedef struct auvpf_handle auvpf_handle;

/* Constructor parameters: the ROS params of auv_pf.__init__ (auv_pf.py:27-56,
 * auv_pf.launch:15-20) plus what the GPU path adds. */
typedef struct auvpf_config {
    int32_t struct_size;     /* sizeof(auvpf_config), for forward compatibility      */
    int32_t device;          /* CUDA device ordinal                                  */
    int64_t particle_count;  /* M, ~particle_count (auv_pf.py:27)                    */
    int32_t n_filters;       /* F independent filters sharing map/beams (default 1)  */
    int32_t coarse_level;    /* max-mip level used for ray skipping; -1 = automatic  */
    int64_t global_offset;   /* id of this shard's first particle (multi-GPU), else 0 */
    int64_t global_count;    /* particles per filter over all shards; 0 = M          */
    double motion_cov[6];    /* ~motion_covariance (auv_pf.py:40-44)                 */
    double meas_std;         /* ~measurement_std (auv_pf.py:39)                      */
    double m2o[16];          /* map<-odom 4x4 row-major (auv_pf.py:79-81)            */
    uint64_t seed;           /* Philox key                                           */
    double range_min;        /* measured beams outside [range_min, range_max] or     */
    double range_max;        /*   non-finite are skipped; a raycast MISS = range_max */
} auvpf_config;

/* -- lifecycle ------------------------------------------------------------ */
int auvpf_version(void);
const char *auvpf_last_error(void);
/* auv_pf.__init__ (auv_pf.py:25-119) minus ROS; particles start at pose 0 and
 * are spread by auvpf_init_particles. */
int auvpf_create(const auvpf_config *cfg, auvpf_handle **out);
int auvpf_destroy(auvpf_handle *h);
int auvpf_sync(auvpf_handle *h);
/* the handle's cudaStream_t, for torch.cuda.ExternalStream interop */
void *auvpf_stream(auvpf_handle *h);
int64_t auvpf_particle_count(auvpf_handle *h); /* F*M rows held by this handle (changes with auvpf_set_next_particle_count) */
int64_t auvpf_particle_capacity(auvpf_handle *h); /* ... and what it was created with: the upper bound of the count */

/* -- map and sensor model (no reference counterpart: SURVEY.md section 0) --- */
/* H x W float32 vertex heights, row i at y0 + i*res, column j at x0 + j*res;
 * builds the max-mip pyramid on the device. */
int auvpf_set_dem(auvpf_handle *h, const float *z, int32_t H, int32_t W, double x0, double y0, double res);
/* Map ingestion: rasterise a triangle mesh (n_vertices x 3 float32 map-frame points, n_faces x 3
 * int32 vertex indices) onto the H x W vertex grid of a DEM; where surfaces overlap in plan
 * view the highest wins, vertices no triangle covers get `fill`.  z_out: (H, W) float32, host or
 * device; pass it to auvpf_set_dem.  (The reference has no map format, SURVEY.md section 8f-4;
 * BASELINE config C1 describes its bathymetry as a mesh.) */
int auvpf_rasterise_mesh(auvpf_handle *h, const float *vertices, int64_t n_vertices, const int32_t *faces,
                         int64_t n_faces, int32_t H, int32_t W, double x0, double y0, double res, float fill,
                         float *z_out);
/* B unit beam directions in the sensor (= particle body) frame, (B,3) float32 */
int auvpf_set_beams(auvpf_handle *h, const float *dirs_xyz, int32_t B);

/* -- particle state --------------------------------------------------------- */
/* Particle.__init__ -> add_noise(init_cov) (auv_particle.py:24,30,32-36):
 * pose = mean + sqrt(var) * n.  noise: (F*M, 6) float32 standard normals in
 * the reference's draw order, or NULL for the library's Philox stream. */
int auvpf_init_particles(auvpf_handle *h, const double mean[6], const double var[6], const float *noise);
int auvpf_set_poses(auvpf_handle *h, const double *poses);   /* (F*M, 6) */
int auvpf_get_poses(auvpf_handle *h, double *poses);         /* (F*M, 6) */
int auvpf_get_logw(auvpf_handle *h, float *logw);            /* (F*M)    */
int auvpf_set_logw(auvpf_handle *h, const float *logw);      /* (F*M)    */

/* -- predict: auv_pf.predict (auv_pf.py:213-216) -> Particle.motion_pred
 * (auv_particle.py:38-70).  odom: (F, 11) float64 rows
 * [vx, vy, vz, wz, qx, qy, qz, qw, z, dt, unused]; n_odom == 1 broadcasts one
 * row to every filter.  noise: (F*M, 6) float32 normals (columns 0, 1, 5 are
 * consumed, like the reference) or NULL for Philox. */
int auvpf_predict(auvpf_handle *h, const double *odom, int32_t n_odom, const float *noise);
/* n_steps consecutive predicts (the odometry messages queued between two pings) in one pass
 * over the state: odom is (n_steps, n_odom, 11), noise (n_steps, F*M, 6) or NULL.  Bit-identical
 * to n_steps calls of auvpf_predict. */
#define AUVPF_MAX_PREDICT_STEPS 16
int auvpf_predict_many(auvpf_handle *h, const double *odom, int32_t n_steps, int32_t n_odom, const float *noise);

/* -- update ----------------------------------------------------------------- */
/* auv_pf.update (auv_pf.py:135-167) -> Particle.compute_weight
 * (auv_particle.py:100-106): log N2(gps; p_map[0:2], std^2 I).  gps_xy: (F, 2)
 * in the map frame.  logw_out: (F*M) float32 or NULL (kept on the device). */
int auvpf_update_gps(auvpf_handle *h, const double *gps_xy, float *logw_out);
/* North-star form of compute_weight: raycast every beam from every particle
 * pose (get_p_pose, auv_particle.py:72-83) into the DEM and score
 * logw = -1/(2 std^2) sum_b (r_b - rhat_b)^2 - B_valid log(std sqrt(2 pi)).
 * ranges: (F, B) float32 measured ranges (a LaserScan's ranges[]). */
int auvpf_update_mbes(auvpf_handle *h, const float *ranges, float *logw_out);
/* How the last auvpf_update_mbes treated the particles, (F, 2) int64 per filter: [0] those whose sensor reach
 * leaves the map (second pass of the one-march-per-fan-side kernel, with the map-border handling), [1] those
 * that could not be certified at all (sensor outside the map, centre beam missed, a facet steeper than the fan
 * tilt allows) and were recast ray by ray; synchronises.  A measurement aid (bench.py reports the fractions per
 * cloud shape). */
int auvpf_mbes_flagged(auvpf_handle *h, int64_t *count);
/* Parity probe: expected ranges (n, B) float32 for n explicit poses (n, 6). */
int auvpf_raycast(auvpf_handle *h, const double *poses, int64_t n, float *ranges_out);

/* -- normalise + resample: auv_pf.resample (auv_pf.py:169-192) --------------- */
/* Uses the device-resident log-weights of the last update.  u: F uniforms in
 * [0,1) replacing resampling.py's random() (:157), or NULL for Philox.
 * ancestors_out: (F*M) int32 (filter-local ids) or NULL.  rough_var: 6
 * variances of the post-resample add_noise (auv_pf.py:191-192) fused into the
 * gather, or NULL for none; rough_noise as in auvpf_init_particles.
 * Returns AUVPF_ERR_DEGENERATE (state untouched) when a filter has no finite
 * weight -- reading that flag back is the one host synchronisation of the call;
 * AUVPF_RS_ASYNC skips it. */
int auvpf_resample(auvpf_handle *h, int32_t mode, int32_t flags, const double *u, int32_t *ancestors_out,
                   const double *rough_var, const float *rough_noise);
/* Same, from caller-supplied LINEAR float64 weights (F*M) -- the literal
 * resample(weights) of the reference (normalisation auv_pf.py:172 included). */
int auvpf_resample_weights(auvpf_handle *h, const double *weights, int32_t mode, int32_t flags, const double *u,
                           int32_t *ancestors_out, const double *rough_var, const float *rough_noise);
/* The fixed-point normalised weights of the last resample: (F*M) uint64, w_j = q_j * 2^-52 with
 * q_j = cdf_j - cdf_{j-1}, cdf_j = floor(C_j 2^52 / A), C_j = sum_{i<=j} a_i, a_i = floor(exp(logw_i - max) 2^60):
 * sum_j q_j == 2^52 exactly per filter, and np.cumsum(w) is cdf 2^-52 bit for bit.
 * On a shard of a sharded filter: this shard's slice of the GLOBAL weights. */
int auvpf_get_weights_q(auvpf_handle *h, uint64_t *q);
/* Apply caller-supplied ancestor indices with the reference's in-place slot
 * map (auv_pf.py:183-189 keep/lost/dupes + reassign_poses :195-198). */
int auvpf_apply_indices(auvpf_handle *h, const int32_t *indices, int32_t flags);
/* Particle.add_noise on every particle (auv_pf.py:191-192). */
int auvpf_roughen(auvpf_handle *h, const double var[6], const float *noise);

/* -- estimate: auv_pf.update_loc_pose (auv_pf.py:218-252) -------------------- */
/* mean: (F, 6) [x y z roll pitch yaw_mean_of_wrapped]; cov: (F, 9) row-major
 * 3x3 with the reference's fill pattern ([2,0], [2,1] left 0). */
int auvpf_mean_cov(auvpf_handle *h, double *mean, double *cov);
/* The estimate with options.  flags == AUVPF_EST_REFERENCE is auvpf_mean_cov.
 * AUVPF_EST_WEIGHTED: weighted by the current log-weights (w_i = exp(logw_i - max); meaningful between an
 * update and the resample that consumes it).  AUVPF_EST_CIRCULAR_YAW: mean[5] = atan2(sum w sin yaw,
 * sum w cos yaw) instead of the reference's arithmetic mean of wrapped yaws (:229-231), which is wrong for
 * a cloud that straddles +-pi.  AUVPF_EST_FULL_COV: all nine entries of the symmetric 3x3 covariance
 * (the reference leaves [2,0] and [2,1] zero, :238-246). */
#define AUVPF_EST_REFERENCE 0
#define AUVPF_EST_WEIGHTED 1
#define AUVPF_EST_CIRCULAR_YAW 2
#define AUVPF_EST_FULL_COV 4
int auvpf_estimate(auvpf_handle *h, int32_t flags, double *mean, double *cov);
/* The PoseArray of auv_pf.loc_loop (auv_pf.py:264-285: every particle's position and
 * quaternion_from_euler(roll, pitch, yaw), :272-275) for a STRIDED subsample, built on the device:
 * poses7_out (F, n, 7) float64 rows [x y z qx qy qz qw], particles 0, stride, 2 stride, ... of every
 * filter, n = min(ceil(M / stride), max_poses) returned in *n_out; stride <= 0 picks the stride that yields at
 * most max_poses.  (After a resample every particle has the same weight, so a stride IS a fair subsample; the
 * reference publishes all of them -- 48 MB per 10 Hz tick at 1 M particles.) */
int auvpf_pose_array(auvpf_handle *h, int64_t stride, int64_t max_poses, double *poses7_out, int64_t *n_out);

/* The k particles of LARGEST weight of every filter as Pose entries ("strided / top-k particle
 * subsample" for the PoseArray of loc_loop, auv_pf.py:264-285): poses7_out [F][k][7] = x y z qx qy qz qw, optionally
 * their log-weights [F][k] and particle indices [F][k].  Meaningful between an update and the resample that
 * consumes its weights.  The SET is exact and deterministic -- larger log-weight first, lower index among equal
 * weights, NaN counted as -inf (a radix select on the device, nothing but k rows comes back) -- the ORDER of the
 * rows is not: sort by (log-weight descending, index ascending) if it matters (the Python facade does). */
int auvpf_pose_array_topk(auvpf_handle *h, int64_t k, double *poses7_out, float *logw_out, int32_t *index_out);

/* ---- KLD-style adaptation of the particle count (Fox, "Adapting the sample size in particle
 * filters through KLD-sampling", IJRR 2003).  Not in the reference, whose particle_count is a launch parameter
 * (auv_pf.py:27).
 * auvpf_kld_bins: how many bins of bin_xy x bin_xy [m] x bin_yaw [rad] the current cloud of every filter occupies
 *   (bin index = floor(coordinate * (1 / bin)), yaw shifted by +pi; a hash set on the device; bins_out [F]).
 * auvpf_kld_target: the number of particles that keeps the KL divergence between the sample-based posterior and
 *   the true one below eps with the confidence whose standard-normal quantile is z_quantile, clamped to
 *   [n_min, n_max] (a pure function; same float64 operations as oracle/pf_oracle.py kld_target).
 * auvpf_set_next_particle_count: the NEXT systematic, non-in-place resample of this handle draws n particles
 *   -- positions (u + i) / n over the CDF of the current ones -- and the handle continues with n; 1 <= n <= the
 *   count it was created with, one filter per handle, not particle-sharded.  A resample that is held back by the
 *   N_eff gate, or fails on degenerate weights, leaves the count alone (the request stays pending).  0 cancels. */
int auvpf_kld_bins(auvpf_handle *h, double bin_xy, double bin_yaw, int64_t *bins_out);
int64_t auvpf_kld_target(int64_t bins, double eps, double z_quantile, int64_t n_min, int64_t n_max);
int auvpf_set_next_particle_count(auvpf_handle *h, int64_t n);
/* The same without leaving the device (a sharded filter reduces its ranks' sums with one collective and
 * reads them back once, sharded.py::combine_pose_sums): auvpf_pose_sums_dev enqueues the pass, the 16
 * float64 sums per filter land in the library buffer auvpf_pose_sums_buffer returns -- [sum d_xyz, sum roll,
 * pitch, yaw, wrapped yaw, sum d d^T (xx yy zz xy xz yz), c_xyz], d = position - c, c = the first particle. */
int auvpf_pose_sums_buffer(auvpf_handle *h, void **sums_f64);
int auvpf_pose_sums_dev(auvpf_handle *h);
/* degenerate-weights flag per filter of the last resample, (F) int32; synchronises */
int auvpf_resample_status(auvpf_handle *h, int32_t *degenerate);
/* effective sample size 1 / sum w^2 of the last normalisation, (F) */
int auvpf_neff(auvpf_handle *h, double *neff);
/* N_eff-gated resampling -- the block the reference carries commented out, auv_pf.py:174-181
 * (N_eff = 1/np.sum(np.square(weights)); resample only "if N_eff < self.pc/2.").  ratio = 0
 * (default) resamples on every call like the live reference code; ratio in (0, 1] makes
 * auvpf_resample* hold a filter's resample back -- poses, weights and roughening untouched,
 * identity ancestors -- whenever its N_eff is NOT below ratio * particle_count (the reference's
 * commented value is 0.5).  The decision is taken on the device inside the select kernel (no host
 * round trip on the systematic path) from an exact integer sum of squares, so it is reproducible:
 * N_eff here is auvpf_neff_fixed = (sum_i s_i)^2 / sum_i s_i^2 over s_i = a_i >> 29, a_i the 60-bit integer
 * weights: sums of integers, the same on one GPU or eight.  With AUVPF_RESIDUAL or
 * AUVPF_RS_INPLACE_REF the handle must hold one filter.  A particle-sharded filter is gated on the global sums.
 * A held-back resample still CONSUMES its draws on every path alike (systematic, residual, slot map, sharded):
 * the Philox step counters of the resample uniform and of the roughening advance as if it had happened, so a
 * gated run stays in step with an ungated one that is fed the same explicit draws. */
int auvpf_set_neff_gate(auvpf_handle *h, double ratio);
int auvpf_neff_fixed(auvpf_handle *h, double *neff);          /* (F); synchronises */
/* 1 per filter whose last resample the gate held back, (F) int32; synchronises */
int auvpf_resample_skipped(auvpf_handle *h, int32_t *skipped);

/* -- replay of a recorded sequence -------------------------------- */
/* n_cycles filter cycles enqueued back to back without the host in the loop: cycle c runs the
 * n_steps odometry messages odom[c] ((n_cycles, n_steps, n_odom, 11) float64, HOST memory, rows as
 * in auvpf_predict) in one pass, the MBES update on ranges[c] ((n_cycles, F, B) float32, host or
 * device), the systematic resample with the roughening fused (AUVPF_RS_ASYNC semantics: a filter
 * without any finite weight, or held back by the N_eff gate, carries its state over) and the
 * estimate of auv_pf.update_loc_pose.  The sequence is uploaded once and the estimates -- mean_out
 * (n_cycles, F, 6), cov_out (n_cycles, F, 9), host memory, or both NULL -- are read back once at the
 * end (the call's only synchronisation; without them it returns after enqueueing).  u: (n_cycles, F)
 * uniforms or NULL for Philox.  The result is bit-identical to n_cycles x [auvpf_predict_many,
 * auvpf_update_mbes, auvpf_resample, auvpf_mean_cov]: odom_callback x n_steps + the measurement
 * callback + loc_loop of the reference node (auv_pf.py:201-211,125-133,264-285) per cycle. */
int auvpf_replay(auvpf_handle *h, int32_t n_cycles, int32_t n_steps, int32_t n_odom, const double *odom,
                 const float *ranges, int32_t mode, int32_t flags, const double *u, const double *rough_var,
                 double *mean_out, double *cov_out);

/* -- multi-GPU: ONE filter sharded by particles ------------------------------
 * Rank r of G holds the global particles [r*M, (r+1)*M) (auvpf_config.global_offset / global_count).
 * Only two things cross GPUs (north_star): the global maximum / sum of the weights, and the particles of
 * cross-shard ancestors.  The normalised CDF is a quotient of prefix sums, cdf_j = floor(C_j 2^52 / A)
 * (C_j = sum of the integer weights a_i up to j over ALL shards, A the global total), so a shard's CDF is
 * its local prefix shifted by the integer mass of the shards below it: what travels is 6 integers per rank.
 *
 * Peer-memory path (default): every rank maps the other ranks' slab (state, log-weights, chunk tables and
 * a mailbox) -- auvpf_ipc_export / auvpf_ipc_import between processes (one per GPU; <= 8 ranks of one
 * node), auvpf_peer_attach between handles of one process (several GPUs driven by one process, or several
 * handles on one GPU).  auvpf_shard_resample is then the WHOLE resample of auv_pf.resample
 * (auv_pf.py:169-192) in two kernels and without any host or NCCL call: the weights kernel posts this
 * shard's maximum and, at its end, its totals into every peer's mailbox (payload, system fence, flag);
 * the select kernel waits for the G totals, selects this rank's ancestors in the global CDF by rebuilding
 * the owners' prefix sums from their log-weights, loads the ancestors' state from the owners' memory over
 * NVLink, adds the roughening noise and writes this rank's new state.  Calls must be made by every rank
 * the same number of times.  The N_eff gate (auvpf_set_neff_gate) works on the global sums.  When no shard
 * has any finite weight (or the gate holds the resample back) the state is carried over unchanged;
 * auvpf_shard_degenerate reports the former (that call synchronises).
 * ancestors_out_dev: (M) int64 GLOBAL ancestor ids, device memory, or NULL. */
#define AUVPF_IPC_BLOB_BYTES 256
int auvpf_ipc_export(auvpf_handle *h, void *blob);
int auvpf_ipc_import(auvpf_handle *h, const void *blobs, int32_t n_ranks, int32_t my_rank);
int auvpf_peer_attach(auvpf_handle *h, auvpf_handle *const *handles, int32_t n_ranks, int32_t my_rank);
int auvpf_shard_resample(auvpf_handle *h, double u, const double *rough_var, const float *rough_noise,
                         int64_t *ancestors_out_dev);
int auvpf_shard_degenerate(auvpf_handle *h, int32_t *flag);
/* The same resample for ALL ranks of one process (handles attached with auvpf_peer_attach; handles[r] is rank
 * r): enqueued stage by stage over the ranks, so that it cannot deadlock when the ranks' streams share a GPU
 * or a hardware queue (kernels of different streams are not guaranteed to run concurrently; a rank that
 * called auvpf_shard_resample alone would wait inside its kernel for posts that are queued behind it).
 * ancestors_out_dev: n_ranks device pointers ((M) int64 each) or NULL. */
int auvpf_group_resample(auvpf_handle *const *handles, int32_t n_ranks, double u, const double *rough_var,
                         int64_t *const *ancestors_out_dev);
/* The estimate of the sharded filter (auv_pf.update_loc_pose, auv_pf.py:218-252) without a collective library:
 * every rank's 16 pose sums (k_pose_stats: sums about the shard's first particle, see auvpf_pose_sums_dev) are
 * posted into every peer's mailbox and collected by one small kernel; sums_all_out = host (G, 16) float64, the
 * same table on every rank, to be combined in rank order (smarc_navigation_b200.sharded.combine_pose_sums).
 * Like auvpf_shard_resample: every rank calls it the same number of times.  auvpf_group_pose_sums is the
 * stage-interleaved form for the ranks of one process; sums_all_out = (n_ranks, G, 16), one table per rank. */
int auvpf_shard_pose_sums(auvpf_handle *h, double *sums_all_out);
int auvpf_group_pose_sums(auvpf_handle *const *handles, int32_t n_ranks, double *sums_all_out);
/* Staged path (no peer memory: any topology a collective library reaches).  The caller moves the scalars:
 *   auvpf_shard_max_dev;  all-reduce MAX in place on wmax_i32 (ORDER-PRESERVING int32: an integer MAX is
 *   the float max);  auvpf_shard_weights_dev;  all-gather of totals_u64 (6 uint64: A low, A high, the
 *   three integer sums of the N_eff gate, log-weight buffer id) into (G, 6);  auvpf_shard_emit with the
 *   gathered totals (host or device memory) emits the n_out = out_hi - out_lo offspring whose ancestors
 *   live in this shard -- global output slots [out_lo, out_hi), as 32-byte records {double x, y; float z,
 *   roll, pitch, yaw} and optionally their GLOBAL ancestor ids -- for an all-to-all with contiguous
 *   splits;  auvpf_shard_load_roughen takes the M records this rank received.
 * auvpf_shard_max / auvpf_shard_sum are the same two first stages with the scalars returned to the host
 * (totals_out: 6 uint64). */
int auvpf_shard_buffers(auvpf_handle *h, void **wmax_i32, void **totals_u64);
int auvpf_shard_max_dev(auvpf_handle *h);
int auvpf_shard_weights_dev(auvpf_handle *h);
int auvpf_shard_max(auvpf_handle *h, float *m_out);
int auvpf_shard_sum(auvpf_handle *h, const float *m_global, uint64_t *totals_out);
int auvpf_shard_emit(auvpf_handle *h, double u, const uint64_t *totals_all, int32_t n_ranks, int32_t my_rank, int64_t out_lo,
                     int64_t out_hi, void *records, int64_t *ancestors_out);
/* stage 5: replace the local state by M records received from the exchange */
int auvpf_shard_load(auvpf_handle *h, const void *records, int64_t n);
/* same, with the post-resample add_noise (auv_pf.py:191-192) fused into the load */
int auvpf_shard_load_roughen(auvpf_handle *h, const void *records, int64_t n, const double *rough_var,
                             const float *rough_noise);

/* -- measurement support ----------------------------------------------------- */
int auvpf_profile_enable(auvpf_handle *h, int32_t on);
/* accumulated CUDA-event time [ms] and launch count per kernel class since the
 * last reset; synchronises the stream. */
int auvpf_profile_read(auvpf_handle *h, double ms[AUVPF_K_COUNT], int64_t launches[AUVPF_K_COUNT], int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* AUVPF_H */
