/* rs_b200.h -- C ABI of the B200-native RoboSumo hot path.
 *
 * Drop-in boundary.  Every entry point replaces one interface of the
 * reference; pointers are DEVICE pointers unless the name says "host"; all calls are
 * stream-ordered on the supplied CUDA stream (a `cudaStream_t` passed as void*), return 0 on
 * success or a negative code, and never throw.  `rs_last_error()` describes the last failure.
 *
 *   rs_create / rs_destroy      SubprocVecEnv.__init__/close   subproc_vec_env.py:40-63,83-93
 *                               + gym.make(...)                 robosumo/robosumo/__init__.py:8-105
 *                               + MujocoEnv.__init__            robosumo/robosumo/envs/mujoco_env.py:47-56
 *   rs_reset                    SubprocVecEnv.reset             subproc_vec_env.py:78-82
 *                               -> SumoEnv.reset_model          robosumo/robosumo/envs/sumo.py:232-253
 *   rs_set_state / rs_get_state MujocoEnv.set_state / sim.get_state   mujoco_env.py:110-119,
 *                               mujoco-py/mujoco_py/mjsimstate.pyx:31-39 (qpos|qvel flattening)
 *   rs_step                     SubprocVecEnv.step_async+step_wait    subproc_vec_env.py:65-76
 *                               -> worker auto-reset            subproc_vec_env.py:12-16
 *                               -> sumo_env.SumoEnv.step        sumo_env.py:40-72
 *                               -> SumoEnv._step                sumo.py:120-192
 *                               -> do_simulation / mj_step x5   mujoco_env.py:125-129, mjsim.pyx:115-129
 *   rs_step_host                same call with HOST numpy buffers (what a SubprocVecEnv user holds)
 *   rs_forward_debug            sim.forward() + data.qacc / data.ncon (parity hook)  mjsim.pyx:101-106
 *   rs_policy_*                 PolicyWithValue.step/value/action_probability   policies.py:84-128
 *   rs_vtrace                   Runner.run V-trace block        runner.py:166-200
 *   rs_ppo_*                    PPOModel.train                  model.py:179-213 (graph 51-139)
 */
#ifndef RS_B200_H
#define RS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RS_MAXL 8            /* max legs per agent (spider) */
#define RS_INFO_DIM 8        /* ctrl, lose, win, main, move, push, shaping, flags(bit0 winner, bit1 timeout) */

#define RS_OK 0
#define RS_ERR_ARG -1
#define RS_ERR_CUDA -2
#define RS_ERR_UNSUPPORTED -3

/* status bits (per env, sticky until rs_reset): the analogue of MuJoCo warnings raised as
 * MujocoException by mujoco-py (builder.py:351-369) */
#define RS_STATUS_NAN 1          /* mjWARN_BADQPOS/BADQVEL/BADQACC */
#define RS_STATUS_CONTACT_FULL 2 /* mjWARN_CONTACTFULL: more than RS_MAXCON contacts, extra ones dropped */
#define RS_STATUS_NEWTON_MAXIT 4 /* constraint solver hit its iteration cap */

/* Per-morphology constants in the layout the kernels read (built by morphology.py). */
typedef struct rs_agent_model {
    int L, nq, nv, nu;
    float torso_r, leg_r, armature, damping, gear, adjust_z, reach, pad2;   /* reach: bounding radius about the torso origin */
    float mT, cT[3], IT[9];                 /* torso rigid group: mass, com, inertia about com (torso frame) */
    float iw_torso, iw_aux[RS_MAXL], iw_hip[RS_MAXL], iw_ank[RS_MAXL];   /* body_invweight0 (translational) */
    float iwd_hip[RS_MAXL], iwd_ank[RS_MAXL];                             /* dof_invweight0 */
    float r_hip[RS_MAXL][3], ax_hip[RS_MAXL][3], r_ank[RS_MAXL][3], ax_ank[RS_MAXL][3], e_ank[RS_MAXL][3];
    float m_hip[RS_MAXL], ip_hip[RS_MAXL], ia_hip[RS_MAXL], m_ank[RS_MAXL], ip_ank[RS_MAXL], ia_ank[RS_MAXL];
    float lo_hip[RS_MAXL], hi_hip[RS_MAXL], lo_ank[RS_MAXL], hi_ank[RS_MAXL];
} rs_agent_model;

typedef struct rs_config {
    int num_envs;            /* env pairs on this device */
    int frame_skip;          /* 5   (sumo.py:50) */
    int timestep_limit;      /* 500 (robosumo/__init__.py) */
    int newton_iters;        /* cap on Newton iterations per forward evaluation (default 16; 9 is the most seen in 2.4 M env-steps; MuJoCo: 100) */
    float timestep;          /* 0.01 (tatami.xml:3) */
    float ring_limit;        /* tatami_size + 0.1 (sumo.py:55) */
    float init_pos_noise;    /* 0.1 */
    float init_vel_noise;    /* 0.1 */
    uint64_t seed;           /* Philox key; env e uses stream (seed, e) */
    int device;
    int reserved;
} rs_config;

typedef struct rs_env rs_env;

int rs_agent_model_size(void);
const char* rs_last_error(void);

int rs_create(const rs_config* cfg, const rs_agent_model* agents /* [2], host */, rs_env** out);
void rs_destroy(rs_env* h);
int rs_dims(const rs_env* h, int* nq, int* nv, int* nu, int* obs_a, int* obs_b, int* act_a, int* act_b);

/* mask: uint8[E] (device) or NULL = all.  obs: float[E][obs_a+obs_b] (agent 0 then agent 1). */
int rs_reset(rs_env* h, const uint8_t* mask, float* obs, void* stream);
int rs_set_state(rs_env* h, const float* qpos, const float* qvel, float* obs, void* stream);
int rs_get_state(rs_env* h, float* qpos, float* qvel, int* ep_step, int* status, void* stream);

/* actions: float[E][act_a+act_b]; obs as above; rew float[E][2]; done uint8[E][2];
 * info float[E][2][RS_INFO_DIM]; episode float[E][3] = (r, dr, l) of agent 0, valid where done[e][0]. */
int rs_step(rs_env* h, const float* actions, float* obs, float* rew, uint8_t* done, float* info,
            float* episode, int auto_reset, void* stream);
/* same, HOST buffers; H2D/D2H copies happen inside (pinned staging owned by the handle) */
int rs_step_host(rs_env* h, const float* actions, float* obs, float* rew, uint8_t* done, float* info,
                 float* episode, int auto_reset);

/* parity hook: one forward evaluation at the stored state with the given ctrl (device pointers):
 * qacc float[E][nv], ncon int[E], niter int[E] */
int rs_forward_debug(rs_env* h, const float* ctrl, float* qacc, int* ncon, int* niter, void* stream);


/* ---- learner side (device pointers, stream-ordered) ---------------------------------------------------------------
 * Flat parameter vector = the reference checkpoint order (model.py:153-161 / tf.trainable_variables):
 * pi.fc0.w[D,64] b pi.fc1.w[64,64] b vf.fc0.w[D,64] b vf.fc1.w[64,64] b pi.w[64,A] pi.b[A] logstd[1,A] vf.w[64,1] vf.b[1] */
int rs_param_count(int obs_dim, int act_dim);
/* PolicyWithValue mean / value heads (policies.py:84-128, models.py:93-101): mean [n,A] and/or value [n] (either may be NULL).
 * precision: 0 = FP32 pipe (numerics reference), 1 = tcgen05 tensor cores, tf32 inputs / fp32 accumulate */
int rs_mlp_forward(const float* params, int obs_dim, int act_dim, const float* obs, long long obs_row_stride, int n,
                   float* mean, float* value, int precision, void* stream);
/* up to 4 independent (params, obs) evaluations in ONE launch: the four policy evaluations of a rollout step (runner.py:62-97) */
typedef struct rs_mlp_job { const float* params; const float* obs; long long obs_row_stride; float* mean; float* value;
                            int activation /* 0 relu (policies.py), 1 tanh (policy_zoo/policy.py:52,64) */; int reserved; } rs_mlp_job;
int rs_mlp_forward_multi(const rs_mlp_job* jobs /* host array */, int njobs, int obs_dim, int act_dim, int n, int precision, void* stream);
/* one rollout step of Runner.run (runner.py:62-100): sample a0 ~ pi0(o0), a1 ~ pi1(o1) and the four neglogps */
int rs_rollout_sample(int E, int act_dim, const float* logstd0, const float* logstd1, const float* mu00, const float* mu10,
                      const float* mu11, const float* mu01, unsigned long long seed, unsigned int tick, int deterministic,
                      float* actions, float* nlp0, float* nlp1, float* opp_nlp0, float* opp_nlp1, void* stream);
/* DiagGaussianPd.neglogp (distributions.py:238-241) */
int rs_neglogp(int n, int act_dim, const float* act, const float* mu, const float* logstd, float* out, void* stream);
/* Runner.run's step loop (runner.py:62-104) for a symmetric pair, T steps enqueued by ONE call: per step the four policy
 * evaluations (one launch), the trajectory writes of the observation / done flags, action sampling with the four neglogps, the
 * physics step with auto-reset, and the trajectory writes of the reward terms and episode records -- 5 launches, no host work.
 * All pointers are device pointers.  `obs`, `done` hold the current observation / done flags on entry and the last ones on exit
 * (they are the env's own output buffers).  Trajectory arrays are time-major: mb_obs [2][T][E][D], mb_actions [T][E][2][A],
 * mb_values / mb_nlp / mb_opp_nlp / mb_shaping / mb_main [2][T][E], mb_dones [2][T][E] u8, ep_done [T][E] u8, ep_info [T][E][3];
 * scratch: 4*E*A floats.  Sampling noise is Philox keyed by (seed, tick0 + t). */
typedef struct rs_rollout_io {
    const float* params0; const float* params1;                 /* learner (agent 0) and opponent (agent 1), flat parameter vectors */
    float* obs; float* rew; uint8_t* done; float* info; float* episode;
    float* mb_obs; float* mb_actions; float* mb_values; float* mb_nlp; float* mb_opp_nlp; uint8_t* mb_dones;
    float* mb_shaping; float* mb_main; uint8_t* ep_done; float* ep_info; float* scratch;
} rs_rollout_io;
int rs_rollout(rs_env* h, int T, const rs_rollout_io* io, int precision, unsigned long long seed, unsigned int tick0, int deterministic,
               void* stream);
/* V-trace returns and IS ratios (runner.py:166-200); arrays [2][T][E], last_values [2][E], last_dones [E][2], ratios [3][T][E] */
int rs_vtrace(int T, int E, float gamma, float lam, float rho_bar, float c_bar, const float* rewards, const float* values,
              const uint8_t* dones, const float* nlp, const float* opp_nlp, const float* last_values, const uint8_t* last_dones,
              float* returns, float* ratios, void* stream);
/* PPOModel.train (model.py:179-213), split so that a data-parallel caller can all-reduce between the pieces */
int rs_adv_moments(const int* idx, int n, const float* returns, const float* values, double* sums, void* stream);
long long rs_ppo_workspace_floats(int obs_dim, int act_dim, int max_minibatch);
/* local part of one minibatch (2 launches: tile kernel + deterministic reduction of the per-block partials): grad_stats [P + 4] =
 * sum over the n local samples of d loss / d theta (divided by global_n) followed by the stat sums (pg, vf, approxkl, clipfrac);
 * stats5 (may be NULL): stats5[2] receives the policy entropy of the PRE-update parameters */
int rs_ppo_grad(const float* params, int obs_dim, int act_dim, const float* obs, const float* actions, const float* returns,
                const float* values, const float* old_nlp, const float* weights, const int* idx, int n, long long global_n,
                const double* adv_sums, float cliprange, float ent_coef, float vf_coef, float* workspace, float* grad_stats,
                float* log_ratio, double* stats5, int precision, void* stream);
/* global-norm clip (model.py:130-132) and TF-style Adam (model.py:121,139) on the (all-reduced) gradient in ONE launch (the
 * entropy term is added by rs_ppo_grad: each rank its share n / global_n of it; ent_coef here is unused, kept for the signature); stats5 (may be NULL) receives [pg_loss, vf_loss, ., approxkl, clipfrac] = stat sums / global_n */
int rs_adam_step(float* params, float* m, float* v, float* grad, int obs_dim, int act_dim, float ent_coef, float max_grad_norm,
                 float lr, long long step_t, float beta1, float beta2, float eps, float* gnorm_out, long long global_n, double* stats5,
                 void* stream);
/* DATA-PARALLEL minibatch schedule on the device (alg_ppo.py:355-398 over env shards): from the GLOBAL permutation of one epoch
 * (int32, device; every rank holds the same one) the LOCAL indices of every minibatch that fall in this rank's sample range
 * [lo, hi): out_idx [nmb][nbatch_train] (minibatch m at out_idx + m * nbatch_train, order preserved), counts [nmb] */
int rs_epoch_split(const int* perm, long long n_global, int nbatch_train, long long lo, long long hi, int* out_idx, int* counts, void* stream);
/* the same cut on the HOST (no GPU involved), for the helper thread that draws the permutation: `local` needs hi - lo + 1 slots */
int rs_epoch_split_host(const void* perm, int elem_bytes /* 4 or 8 */, long long n_global, int nbatch_train, long long lo, long long hi, int* local, int* counts);
/* advantage moments (model.py:182-185) of ALL minibatches of an epoch in one launch: sums [nmb][2] doubles; idx [nmb][cap] (NULL =
 * identity over n_total samples), counts [nmb] (NULL = full slices).  Returns and values are constant during an update, so a
 * data-parallel caller all-reduces these once per epoch instead of once per minibatch */
int rs_adv_moments_multi(const int* idx, const int* counts, int nmb, int cap, long long n_total, const float* returns, const float* values,
                         double* sums, void* stream);
/* the five scalars PPOModel.train returns (model.py:137, loss_names): stats5 = [pg_loss, vf_loss, entropy, approxkl, clipfrac] as
 * doubles, from the (all-reduced) stat sums behind the gradient in grad_stats and the logstd of `params`; call it BEFORE
 * rs_adam_step: the reference evaluates the entropy with the pre-update parameters (same session.run as the train op) */
int rs_ppo_stats(const float* grad_stats, const float* params, int obs_dim, int act_dim, long long global_n, double* stats5, void* stream);

/* tcgen05 self-test (debug hook): D[128,64] = op(A)*op(B) with kind::tf32; prm13 = {a_rows,a_cols,b_rows,b_cols,a_mn,b_mn,
 * a_lbo,a_sbo,a_step,b_lbo,b_sbo,b_step,nk} (bytes), host pointer */
int rs_tc_selftest(const int* prm13, const float* A, const float* B, float* D, void* stream);

/* OR of the status bits any env raised since the last clear (out2_host[0]) and the number of env-steps that raised one
 * (out2_host[1]); HOST pointer, synchronises with `stream`.  Auto-reset clears an env's own status word but not this latch:
 * check it once per rollout, where mujoco-py's warning callback would have raised inside the worker (builder.py:351-369) */
int rs_status_latch(rs_env* h, int* out2_host, int clear, void* stream);
/* env.seed(s) (run.py:73-83, mujoco_env.py:82-84): re-keys the Philox streams of the reset states; env e keeps stream (seed, e) */
int rs_seed(rs_env* h, unsigned long long seed);

/* diagnostics of the last rs_step: int[E][4] = (Newton iterations, evaluations with an inter-agent contact, contacts) summed over
 * the 20 forward evaluations of the step, and the largest iteration count of a single evaluation (cf. mjData.solver_iter / ncon) */
int rs_get_diag(rs_env* h, int* diag, void* stream);

/* ---- gradient all-reduce over NVLink peer memory (one node, one process per GPU) -----------------------------------------------
 * Replaces the torch.distributed / NCCL all-reduce of the flat gradient(+stats) inside a minibatch step (the pattern of
 * baselines/baselines/common/mpi_adam_optimizer.py:21-46) with ONE kernel on the learner's own stream: every rank owns a
 * double-buffered symmetric buffer that its peers map through CUDA IPC; rs_ppo_grad writes the local gradient into
 * rs_peer_send_buffer(), rs_peer_allreduce then (1) raises this rank's step counter in every peer's flag array, (2) waits for the
 * peers' counters and (3) sums the `nfloats` entries over the ranks in rank order (bit-identical on every rank) into `out`.
 * No host synchronisation and no second stream; a peer that never arrives makes the kernel give up after ~9 s and latches an
 * error that rs_peer_error() reports.  Handles are exchanged by the caller (dist.py: one all_gather of rs_peer_handle_bytes()). */
typedef struct rs_peer rs_peer;
int rs_peer_create(int rank, int world, long long nfloats, int device, rs_peer** out);
int rs_peer_handle_bytes(void);
int rs_peer_export(rs_peer* p, void* handle_out);
int rs_peer_connect(rs_peer* p, const void* all_handles);          /* [world][rs_peer_handle_bytes()] in rank order */
float* rs_peer_send_buffer(rs_peer* p);                            /* where the NEXT rs_peer_allreduce expects this rank's data */
int rs_peer_allreduce(rs_peer* p, float* out, long long nfloats, void* stream);
int rs_peer_error(rs_peer* p);                                     /* non-zero after a timed-out wait (synchronises the device) */
void rs_peer_destroy(rs_peer* p);

/* data-parallel minibatch step in ONE call: rs_ppo_grad (into the peer buffer) -> rs_peer_allreduce -> rs_adam_step, 4 launches and no
 * host work in between, so that the host stays ahead of the GPU (with three separate calls the Python side of a minibatch took as
 * long as its kernels and every hiccup of one of N ranks stalled all of them).  adv_sums: GLOBAL moments of the minibatch. */
int rs_ppo_minibatch_step_peer(rs_peer* peer, float* params, float* m, float* v, int obs_dim, int act_dim, const float* obs, const float* actions,
                               const float* returns, const float* values, const float* old_nlp, const float* weights, const int* idx, int n,
                               long long global_n, float cliprange, float ent_coef, float vf_coef, float max_grad_norm, float lr, long long step_t,
                               float* workspace, float* grad_stats, const double* adv_sums, float* gnorm_out, double* stats5, float* log_ratio,
                               int precision, void* stream);
/* single-GPU convenience: [rs_adv_moments ->] rs_ppo_grad -> rs_adam_step in ONE call, 3 launches (a data-parallel caller uses the
 * pieces and all-reduces grad_stats between them).  beta1 = 0.9, beta2 = 0.999, eps = 1e-5 as model.py:121; adv_sums [2] doubles
 * (already filled when moments_ready != 0, e.g. by rs_adv_moments_multi) and grad_stats [P + 8] floats are caller-owned device buffers. */
int rs_ppo_minibatch_step(float* params, float* m, float* v, int obs_dim, int act_dim, const float* obs, const float* actions,
                          const float* returns, const float* values, const float* old_nlp, const float* weights, const int* idx, int n,
                          float cliprange, float ent_coef, float vf_coef, float max_grad_norm, float lr, long long step_t,
                          float* workspace, float* grad_stats, double* adv_sums, int moments_ready, float* gnorm_out, double* stats5,
                          float* log_ratio, int precision, void* stream);

/* HOST function (no GPU): the reference's per-epoch minibatch permutation, `np.random.shuffle(inds)` on the legacy global
 * RandomState (alg_ppo.py:364), replayed bit-exactly (swap partners drawn a block ahead and prefetched: on par with NumPy at 0.5 M indices, 2.3x faster at 8 M).  `key` [624] and
 * `*pos` are the MT19937 state as returned by np.random.get_state() and are advanced in place; `x` [n] is permuted in place.
 * Algorithm restated from NumPy's legacy generator (RandomState.shuffle -> _shuffle_raw: for i = n-1 .. 1: j = interval(i); swap;
 * interval = masked rejection sampling on 32-bit MT19937 outputs while i <= 0xffffffff). */
int rs_legacy_shuffle(uint32_t* key, int* pos, int64_t* x, long long n);
/* the same permutation and generator state on an int32 array (n < 2^31): half the bytes under the random accesses */
int rs_legacy_shuffle32(unsigned int* key, int* pos, int* x, long long n);

/* number of kernels launched by this library since load (bench.py's gpu_launches) */
long long rs_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
