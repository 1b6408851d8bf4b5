/* libreal_b200.so — C ABI of the B200-native batched REALRobot2020 simulator.
 *
 * The reference (AIcrowd/real_robots) has no FFI of its own: its hot path is a sequence of pybullet C-API
 * calls issued from Python.  Each entry point below replaces the group of reference calls
 * cited beside it (paths relative to the reference repo).  All array pointers are DEVICE pointers owned by
 * the caller unless the name ends in _host; a ctx is bound to one CUDA device, is not thread-safe and never
 * allocates after rr_create.  Work is enqueued on the `stream` argument (a cudaStream_t passed as void*;
 * NULL = legacy default stream); nothing synchronises implicitly except the *_host helpers.
 * Return value: 0 on success, <0 on error (rr_last_error gives the text).
 *
 * Layouts: N = n_envs.  Batched tensors are row-major [N, ...] exactly as listed.
 */
#ifndef REAL_B200_H
#define REAL_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define RR_NUM_JOINTS 9        /* robot.py:26  action / joint_positions length */
#define RR_NUM_DOF 11          /* movable joints of kuka_gripper.urdf (IK output length) */
#define RR_NUM_TOUCH 4         /* robot.py:29 */
#define RR_EYE_W 320           /* robot.py:30 */
#define RR_EYE_H 240           /* robot.py:31 */
#define RR_MAX_OBJECTS 3       /* robot.py:49-50 cube, tomato, mustard */
#define RR_NUM_LINKS 16        /* robot links excluding the fixed base */

typedef struct rr_ctx rr_ctx;

/* Physics / solver parameters (pybullet world defaults after Kuka.reset()'s resetSimulation). */
typedef struct rr_config {
    double dt;                  /* env.py:202-204  0.005 */
    int    solver_iterations;   /* 50 */
    double erp;                 /* 0.2 */
    double erp2;                /* 0.08 */
    double linear_slop;         /* 1e-5 */
    double warmstart;           /* 0.1 */
    double residual_threshold;  /* 1e-7 */
    double lin_damping, ang_damping;   /* 0.04, 0.04 */
    double max_coord_vel;       /* 100 */
    double motor_kp, motor_kd, motor_max_force;   /* 0.1, 1.0, 1e5 */
    double limit_max_impulse;   /* 100 */
    int    use_box_inertia;     /* 1 */
    double gravity;             /* env.py:208  -9.81 */
    int    enable_contacts;     /* 1 */
    double contact_breaking;    /* 0.02 */
    double margin;              /* 0.001 */
    double restitution_threshold; /* 0.2 */
    double friction_erp;        /* 0.2 */
    int    torsional_friction;  /* 1: spinning / rolling friction rows (urdf/cube.urdf:6-7, kuka_gripper.urdf:294-295 ...) */
} rr_config;

void rr_config_default(rr_config* cfg);
int  rr_config_size(void);

/* replaces: gym.make -> REALRobotEnv.__init__ (env.py:36-122) + the URDF/OBJ loading inside reset
 * (robot.py:120-123,165-185,214-226): assets are read once from the compiled blob.
 * Several contexts may be alive on one device (e.g. a 1-object and a 3-object scene): the device-wide constant tables are
 * re-bound, with a device synchronisation, whenever a call arrives for a context other than the one that used them last. */
int  rr_create(rr_ctx** out, int device, int n_envs, int n_objects, const char* asset_blob_path, const rr_config* cfg);
int  rr_destroy(rr_ctx* ctx);
const char* rr_last_error(rr_ctx* ctx);   /* ctx may be NULL: error of the last failed rr_create */
int  rr_n_envs(rr_ctx* ctx);
int  rr_n_objects(rr_ctx* ctx);

/* replaces: REALRobotEnv.reset (env.py:206-219) -> Kuka.reset / robot_specific_reset (robot.py:120-123,165-185).
 * env_mask: u8[N] (non-zero = reset this env) or NULL = all. */
int  rr_reset(rr_ctx* ctx, const uint8_t* env_mask, void* stream);

/* replaces: BodyPart.reset_pose on an object (env.py:161 set_goal; robot.py:125-129 reset_object).
 * pos_quat: f64[N,7] xyz + quaternion xyzw; velocities are zeroed (resetBasePositionAndOrientation). */
int  rr_set_object_pose(rr_ctx* ctx, const uint8_t* env_mask, int obj, const double* pos_quat, void* stream);

/* test / checkpoint hooks: full state in and out.  q,qd: f64[N,11]; obj: f64[N,3,13] = pos3 quat4 lin3 ang3. */
int  rr_set_state(rr_ctx* ctx, const double* q, const double* qd, const double* obj, void* stream);
int  rr_get_state(rr_ctx* ctx, double* q, double* qd, double* obj, void* stream);

/* replaces: step_joints up to and including scene.global_step() (env.py:326-340):
 * limitActionByJoint (env.py:314-321), control_objects_limits (env.py:257-264), Kuka.apply_action
 * (robot.py:188-201), one pybullet stepSimulation at dt (env.py:340).  joint_cmd: f64[N,9].
 * Stream semantics: everything the caller can observe is ordered on `stream`.  Internally the context owns two side streams: when
 * rr_step returns, the solver kernel of the few envs with very long constraint schedules and the collision detection of the NEXT
 * step (which only reads the new poses) are still in flight there.  Every other entry point first makes its stream wait for them;
 * rr_render (eye camera, all envs) does not — its kernel takes those envs last and waits for each one on the device — so calling
 * rr_render right after rr_step overlaps the three.  The split of the solver only pays when a render follows (back-to-back physics
 * steps are faster unsplit), so a step is split iff the previous step was followed by an rr_render of all envs; RR_HEAVY_SOLVER=1 / 0
 * (read at rr_create) forces it on / off, RR_EAGER_COLLIDE=0 moves the collision detection back to the beginning of its own step.
 * Same results in every combination; with both at 0 the kernels of a step are strictly sequential on `stream` (needed to capture
 * a step into a CUDA graph). */
int  rr_step(rr_ctx* ctx, const double* joint_cmd, void* stream);

/* replaces: Kuka.calc_state (robot.py:203-211), get_touch_sensors (robot.py:152-163) and, for R1,
 * get_all_used_objects (env.py:240-244).  joint_pos f64[N,9]; touch f64[N,4]; obj_pos_quat f64[N,n_objects,7] or NULL. */
int  rr_read_state(rr_ctx* ctx, double* joint_pos, double* touch, double* obj_pos_quat, void* stream);

/* replaces: get_retina -> EyeCamera.renderTarget -> getCameraImage/TinyRenderer (env.py:249-255,536-567).
 * env_mask u8[N] = per-env 'render' flag (NULL = all); masked-out envs get the reference's cached zero
 * arrays (env.py:114-122,273-275).  rgb u8[N,240,320,3]; depth f32[N,240,320]; mask i32[N,240,320] or NULL.
 * The table and the robot base (fixed bodies seen by a fixed camera) are rasterised once in rr_create and every env's
 * z-buffer starts from that layer; the images are bit-identical to rasterising every shape (what rr_render_frames and
 * rr_render_camera do, and what rr_render does when the environment variable RR_STATIC_LAYER=0 is set at rr_create). */
int  rr_render(rr_ctx* ctx, const uint8_t* env_mask, uint8_t* rgb, float* depth, int32_t* mask, void* stream);

/* replaces: REALRobotEnv.render('rgb_array') -> EnvCamera.render (env.py:221-228, 470-513): same rasteriser with a
 * caller-supplied camera.  view16_host / proj16_host: column-major float[16] HOST arrays as pybullet's
 * computeViewMatrix* / computeProjectionMatrixFOV return them.  Synchronises the stream. */
int  rr_render_camera(rr_ctx* ctx, const float* view16_host, const float* proj16_host, const uint8_t* env_mask,
                      uint8_t* rgb, float* depth, int32_t* mask, void* stream);

/* test hook: render from explicit body frames instead of the simulated state.
 * frames f64[N,22,12] = world_from_local rotation (row-major 9) + position (3) per scene shape. */
int  rr_render_frames(rr_ctx* ctx, const double* frames, uint8_t* rgb, float* depth, int32_t* mask, void* stream);
int  rr_shape_frames(rr_ctx* ctx, double* frames /* f64[N,22,12] */, void* stream);

/* replaces: pybullet.calculateInverseKinematics(0, 7, pos, orn, maxNumIterations, residualThreshold)
 * (env.py:372-375, 423-426), seeded from each env's current joint state.  pos_quat f64[N,7]; joints_out f64[N,11]. */
int  rr_ik(rr_ctx* ctx, const uint8_t* env_mask, const double* pos_quat, double* joints_out, int max_iter, double residual, void* stream);

/* replaces: BodyPart.get_pose() of a robot link = getLinkState()[0:2], the COM frame (env.py:230-232).
 * link_index -1 = base .. 15; pos_quat f64[N,7]. */
int  rr_link_pose(rr_ctx* ctx, int link_index, double* pos_quat, void* stream);

/* replaces: evaluateGoal (env.py:181-200): score[n] = sum_obj exp(-(ln4/0.10) * |p - goal|).
 * goal_xyz f64[N,n_objects,3]; score f64[N]. */
int  rr_score(rr_ctx* ctx, const double* goal_xyz, double* score, void* stream);

/* Host-buffer convenience path (what the N=1 gym facade and bench.py's e2e leg use): copies the command in,
 * steps, optionally renders, and copies every observation back; synchronises the stream before returning.
 * All pointers are HOST pointers; any output pointer may be NULL. render_mask_host u8[N] or NULL (= no render). */
int  rr_step_host(rr_ctx* ctx, const double* joint_cmd_host, const uint8_t* render_mask_host,
                  double* joint_pos_host, double* touch_host, double* obj_pos_quat_host,
                  uint8_t* rgb_host, float* depth_host, int32_t* mask_host);
/* The same step, split: _begin enqueues it (commands in, physics, read-out, render, every copy back into the given HOST
 * buffers) and returns at once; _end waits for the oldest step begun.  Up to two steps may be in flight (device staging is
 * double buffered), so an open-loop driver — bench.py's random policy, a replayed plan — overlaps the physics of step t+1
 * with the image copies of step t; the host buffers of a step must stay untouched until its _end returns.  rr_step_host is
 * _begin followed by _end.  (Not in the reference, whose step is synchronous: env.py:326-356.) */
int  rr_step_host_begin(rr_ctx* ctx, const double* joint_cmd_host, const uint8_t* render_mask_host,
                        double* joint_pos_host, double* touch_host, double* obj_pos_quat_host,
                        uint8_t* rgb_host, float* depth_host, int32_t* mask_host);
int  rr_step_host_end(rr_ctx* ctx);

/* Number of kernels this library has launched on ctx since creation (bench.py's gpu_launches). */
long long rr_launch_count(rr_ctx* ctx);

/* Optional per-kernel device timing (CUDA events on the launch stream) for bench.py's roofline line.
 * rr_profile(ctx, 1) starts recording, rr_profile(ctx, 0) stops; rr_profile_read synchronises the device and
 * returns, per kernel kind (0 = physics step = k_step_dyn + k_step_solve, 1 = render set-up, 2 = rasteriser, 3 = collision), the summed
 * duration in ms and the number of launches since profiling was enabled, then clears the counters. */
#define RR_PROFILE_KINDS 4
int  rr_profile(rr_ctx* ctx, int enable);
int  rr_profile_read(rr_ctx* ctx, double* ms_total /*[4]*/, long long* launches /*[4]*/);

/* replaces: getContactPoints as used by Kuka.get_contacts (robot.py:131-150): the cached contact manifolds of the
 * last step.  hdr i32[N,32,2] = (candidate pair index or -1, point count); pts f64[N,32,4,17] = per point
 * localA3 localB3 normalOnB3 distance normalImpulse worldA3 worldB3; pairs i32[96,2] = (shapeA, shapeB) per pair index
 * (HOST pointer, may be NULL); returns the number of candidate pairs through n_pairs_host (may be NULL). */
#define RR_MAX_MANIFOLDS_PER_ENV 32
#define RR_CONTACT_POINT_DOUBLES 17
int  rr_contacts(rr_ctx* ctx, int32_t* hdr, double* pts, int32_t* pairs_host, int* n_pairs_host, void* stream);

/* checkpoint hook: write the persistent contact manifolds back (same layout as rr_contacts returns them).  The manifolds —
 * contact points with their accumulated normal impulses, i.e. the solver's warm start — are the only state of the path besides
 * (q, qd, obj); rr_set_state clears them, so a resume is rr_set_state followed by rr_set_contacts and then continues bit for
 * bit.  (Not in the reference, which has no checkpoint of the simulation: SURVEY section 5.) */
int  rr_set_contacts(rr_ctx* ctx, const int32_t* hdr, const double* pts, void* stream);

/* Per-env diagnostics of the last step: i32[N,4] = rows, iterations, contacts, flags (bit 0: the 32-contact cap was hit,
 * bit 4: a non-empty contact manifold found none of the 32 manifold slots free and was dropped,
 * bit 1: every contact was object-vs-static, so the solver took its object-parallel path; bit 2: a NaN / infinite joint
 * command arrived — where the reference asserts np.isfinite(a).all() (robot.py:189) a batched step cannot raise per env, the
 * command acts as a full +-max step and never reaches the state; bit 3: the state had become non-finite and the env was put
 * back into its initial state, as rr_reset would; bits 16-31: the length of the env's Gauss-Seidel schedule in slots per iteration,
 * a friction-pair slot counted twice — what orders the envs inside the solver kernel). */
int  rr_diag(rr_ctx* ctx, int32_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
