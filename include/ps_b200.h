/* ps_b200.h — C ABI of libps_b200.so: the B200-native batched multi-world stepper
 * that replaces the body of PhysicsSimulatorCore's per-step hot path.
 *
 * Citations are relative to /root/reference/PhysicsSimulatorCore/.
 *
 * Drop-in boundary: the one reference call site that changes is
 *   Simulator.updateSimulation (Simulator.fs:47-62):
 *       simulatorState |> resolveCollisions |> SimulatorState.update dt
 *   becomes  marshal -> ps_batch_step -> unmarshal.
 * Everything the F# side keeps computing (RigidBodyPrototype.build: inertia, R from
 * yaw/pitch/roll, gravity force per body — RigidBodyPrototype.fs:69-90,
 * SimulatorState.fs:26-36,67-83) arrives here already evaluated.
 *
 * Conventions
 *  - plain pointers + sizes, FP64 everywhere, no callbacks, no CPU fallback;
 *  - host buffers are borrowed for the duration of a call only;
 *  - every function returns 0 or a negative ps_status; text via ps_last_error()
 *    (thread-local).  The F# shim maps the classes back to the reference's
 *    exception types;
 *  - a handle is bound to ONE CUDA device and is not internally locked: the
 *    reference serialises all access with SemaphoreSlim(1) (Simulator.fs:34);
 *    distinct handles may be used concurrently.  Multi-GPU = one process/handle
 *    per device over a contiguous range of worlds.
 *  - per-body arrays are "body index fastest inside a world, worlds concatenated";
 *    vectors are xyz-interleaved (3 doubles per body), orientation is ROW-MAJOR
 *    9 doubles per body; ids are the index inside the world
 *    (SimulatorState.fs:41-42); static <=> mass == +inf (Entities/Base.fs:13-20).
 */
#ifndef PS_B200_H
#define PS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    PS_OK = 0,
    PS_ERR_INVALID_ARGUMENT = -1, /* ArgumentException  (invalidArg: RigidBodyPrototype.fs:61-67, RigidBody.fs:58-65) */
    PS_ERR_OUT_OF_RANGE = -2,     /* KeyNotFoundException (bad world/body id: SimulatorState.fs:64,113-129) */
    PS_ERR_CUDA = -3,             /* no reference analogue: device failure, never a silent reroute */
    PS_ERR_NAN = -4,              /* NaN/Inf detected in a world (Appendix B Q20) */
    PS_ERR_INVALID_OPERATION = -5,/* InvalidOperationException (invalidOp: Simulator.fs:124-125,148-149,158) */
    PS_ERR_CAPACITY = -6          /* a caller-provided output buffer is too small */
} ps_status;

/* StepConfiguration (Configuration.fs:28-55) + Configuration.SimulationStepInterval
 * (Simulator.fs:16-27), flattened. */
typedef struct {
    double epsilon;                 /* Epsilon = 1e-4 */
    double baumgarte;               /* BaumgarteTerm = 0.2 */
    double allowed_penetration;     /* AllowedPenetration = 0.001 */
    double max_correction_velocity; /* MaxCorrectionVelocity = 1.0 */
    double dt;                      /* SimulationStepInterval.TotalSeconds = 0.01 */
    int32_t solver_iterations;      /* CollisionSolverIterationCount = 10 */
    int32_t enable_friction;        /* EnableFriction = true */
    int32_t integrator_kind;        /* 0 FirstOrder, 1 AugmentedSecondOrder (Configuration.fs:5-7) */
    int32_t broadphase_kind;        /* 0 Dummy: all (i<j), lexicographic (BroadPhase.fs:33-39).
                                     * 1 SpatialTree (BroadPhase.fs:100-127): the persistent 2^3-tree of
                                     *   Utilities/SpatialTree.fs decides the candidates and their ORDER; it lives on
                                     *   the host (one per world, csrc/ps_octree.hpp), is fed with the bodies' extents
                                     *   from the device before every step and hands the ordered list back — one
                                     *   host round trip per step, meant for the Gui-scene sized worlds that select
                                     *   this mode; the device work per pair is the same as in mode 0 */
    int32_t restore_joints;         /* 0 = reference default (Simulator.fs:55 is commented out) */
    int32_t record_contacts;        /* 1: keep the last step's Collisions map for ps_batch_get_contacts */
    int32_t exact_all_pairs;        /* 1: parity/debug — disable the conservative cull, walk all pairs serially */
    /* SpatialTreeConfiguration (Configuration.fs:13-23), read when broadphase_kind == 1 */
    int32_t tree_leaf_capacity;     /* LeafCapacity = 4 */
    int32_t tree_max_depth;         /* MaxDepth = 10 */
    int32_t aabb_mode;              /* extents handed to the SpatialTree: 0 = reference (a sphere's box has side =
                                     * radius, SimulatorObject.fs:58-71, Appendix B Q5), 1 = conservative (side = 2r).
                                     * Only read when broadphase_kind == 1: the Dummy order tests every pair. */
    double tree_min[3];             /* SpaceBoundaries.Min = (-15,-15,-15) */
    double tree_max[3];             /* SpaceBoundaries.Max = ( 15, 15, 15) */
} ps_step_config;

/* gravity is NOT in the config: the F# side folds it into per-body force
 * (SimulatorState.fs:26-36,67-83), so AddObject's hard-coded earth constants (Q21)
 * fall out for free. */

typedef struct {
    int32_t n_worlds;
    const int32_t* world_offset; /* n_worlds+1 prefix offsets into the body arrays */
    const int32_t* shape;        /* 0 sphere, 1 box          (Entities/Collider.fs:5-7) */
    const double* dims;          /* 3/body: (radius,0,0) | (XSize,YSize,ZSize) */
    const double* mass;          /* +inf = Mass.Infinite */
    const double* inertia_inv;   /* 3/body: diagonal of PrincipalRotationalInertiaInverse (RigidBody.fs:76) */
    const double* pos;           /* 3/body */
    const double* vel;           /* 3/body */
    const double* rot;           /* 9/body row-major */
    const double* ang_mom;       /* 3/body angular MOMENTUM L (RigidBody.fs:6-8) */
    const double* force;         /* 3/body ExternalForces  (SimulatorState.fs:13-15) */
    const double* torque;        /* 3/body ExternalTorques */
    const double* elasticity;    /* ElasticityCoeff */
    const double* mu_kinetic;    /* KineticFrictionCoeff */
    const double* mu_static;     /* StaticFrictionCoeff (round-tripped only, Q7) */
} ps_bodies_view;

/* ball-socket joints (Joints.fs:10-53); anchors are LOCAL offsets R^T (p_world - x) */
typedef struct {
    int32_t n_joints;
    const int32_t* world;
    const int32_t* body_a;
    const int32_t* body_b;
    const double* anchor_a; /* 3/joint */
    const double* anchor_b; /* 3/joint */
} ps_joints_view;

/* per-batch counters, summed over worlds and over all steps since create/reset */
typedef struct {
    int64_t steps;            /* world-steps executed */
    int64_t body_steps;       /* dynamic bodies x steps */
    int64_t pairs_tested;     /* candidate pairs that reached the narrow phase */
    int64_t pairs_colliding;  /* |HIT| summed */
    int64_t contacts;         /* contact points summed */
    int64_t fallback_steps;   /* world-steps that restarted (wider margin) or took the exact serial walk */
    int64_t nan_worlds;       /* worlds whose state holds a NaN/Inf now */
    int64_t contact_overflow; /* manifolds truncated at PS_MAX_CONTACTS (must stay 0) */
    int64_t reference_exceptions; /* pair tests on which the reference would have thrown (Option.get of None,
                                     Seq.find / List.maxBy on nothing); treated as "no collision" here */
    double max_penetration;   /* max |penetration| seen */
    /* why a world-step left the wavefront path (sum = fallback_steps) */
    int64_t fallback_margin_steps;   /* a culled pair could not be ruled out any more: retried with an 8x margin */
    int64_t fallback_capacity_steps; /* near list or level table full: exact serial walk */
    int64_t fallback_static_steps;   /* a static body's pose still moved under integrate(): exact serial walk */
    /* world-steps in which a body left its motion margin and its culled pairs were re-examined with the grown
     * sphere; only those where a culled pair could no longer be ruled out count as fallback_margin_steps */
    int64_t margin_verified_steps;
} ps_stats;

#define PS_MAX_CONTACTS 20 /* worst case of the reference clipper is 19 */

/* Contact-feature id (the reference stores none, CollisionDetection.fs:9-23): the tuple
 * of discrete decisions the reference takes to emit a contact, packed in 32 bits.
 *  [1:0]  pair kind: 0 sphere-sphere, 1 sphere-box (sphere first), 2 box-sphere, 3 box-box
 *  sphere/box:  [6:4] collided face index 0..5 (Box.fs:42-49 order)
 *  box-box:     [3:2] origin 0 Face1, 1 Face2, 2 Edge (SAT.fs:14-17)
 *     face:  [6:4] reference face, [9:7] incident face, [14:10] ordinal in the emitted list,
 *            [22:15] provenance of the clipped vertex: bit7=0 -> original incident vertex [1:0];
 *                    bit7=1 -> intersection made by clip plane [5:4] on polygon edge [3:0]
 *     edge:  [6:4] axis i of box 1, [9:7] axis j of box 2, [14:10]/[19:15] supporting directed
 *            edge index (0..23, face-major order of OrientedBox.getEdges) on box 1 / box 2
 */

/* device selection for handles created afterwards by this thread; < 0 keeps the current device */
int ps_init(int32_t device);
const char* ps_last_error(void);
int ps_version(void);

/* SimulatorStateBuilder.fromPrototypes/withConfiguration (SimulatorState.fs:38-83) */
int ps_batch_create(void** handle, const ps_bodies_view* bodies, const ps_joints_view* joints,
                    const ps_step_config* config);
int ps_batch_destroy(void* handle);

/* Simulator.updateSimulation x n_steps (Simulator.fs:47-62): resolveAll then update, per world */
int ps_batch_step(void* handle, int32_t n_steps);
/* same, but on a caller stream (cudaStream_t passed as void*) and without the final synchronise.  The call orders
 * the caller's stream after everything the handle's own stream holds, and the handle's stream (so
 * ps_batch_synchronize and every later call on the handle) after the work launched here.  CUDA-graph capture of the
 * caller's stream: supported for every Dummy-order batch — the fused kernel is one launch, the batch-wide path counts its
 * rounds on the device (no host round trip inside a step); a SpatialTree-order batch (one host round trip per step for
 * the tree) returns PS_ERR_INVALID_OPERATION while capturing.  During a capture the two orderings above are the
 * capture's business. */
int ps_batch_step_async(void* handle, int32_t n_steps, void* cuda_stream);
int ps_batch_synchronize(void* handle);

/* Simulator.AllPhysicalObjects read-back (Simulator.fs:81-82): worlds [w0,w1), layout as the view */
int ps_batch_get_state(void* handle, int32_t w0, int32_t w1, double* pos, double* vel, double* rot,
                       double* ang_mom);
int ps_batch_set_state(void* handle, int32_t w0, int32_t w1, const double* pos, const double* vel,
                       const double* rot, const double* ang_mom);
/* marshal -> n x Simulator.updateSimulation -> unmarshal in ONE call (Simulator.fs:47-62 seen from a caller whose
 * SimulatorState lives in managed memory): equivalent to ps_batch_set_state(all) + ps_batch_step(n) +
 * ps_batch_get_state(all), but pipelined over chunks of worlds on separate streams so that the host<->device copies
 * of one chunk hide behind the step of the others (the per-world kernel over a world range; on the batch-wide path every
 * chunk runs its own chain of rounds with its own counters and a slice of the batch's lists).  Pinned host buffers make the
 * copies asynchronous; pageable ones work, unpipelined.  in and out may be the same arrays. */
int ps_batch_step_host(void* handle, int32_t n_steps, const double* pos_in, const double* vel_in,
                       const double* rot_in, const double* ang_mom_in, double* pos_out, double* vel_out,
                       double* rot_out, double* ang_mom_out);
/* same read-back into caller-owned DEVICE buffers (plain device pointers): no host round trip */
int ps_batch_get_state_device(void* handle, int32_t w0, int32_t w1, double* d_pos, double* d_vel,
                              double* d_rot, double* d_ang_mom, void* cuda_stream);

/* Simulator.AllCollisions / CollisionsIdentifiers of the LAST step (Simulator.fs:73,84-87;
 * CollisionResolver.fs:33-34): pairs in ascending (i,j); contacts computed on the predicted,
 * pre-resolution poses (Q18).  Needs record_contacts=1. */
int ps_batch_get_contacts(void* handle, int32_t world, int32_t* n_pairs, int32_t* pair_ids /*2/pair*/,
                          int32_t* first_contact /*n_pairs+1*/, double* normal, double* position,
                          double* penetration, uint32_t* feature_id, int32_t pair_capacity,
                          int32_t contact_capacity);
/* ordered candidate list the last step walked for `world` (parity/debug): pairs (i<j), >=1 dynamic */
int ps_batch_get_candidates(void* handle, int32_t world, int32_t* n_pairs, int32_t* pair_ids,
                            int32_t pair_capacity);

/* Simulator.ApplyImpulse (Simulator.fs:101-111 -> RigidBody.fs:194-197) */
int ps_batch_apply_impulse(void* handle, int32_t world, int32_t body, const double impulse[3],
                           const double offset[3]);
/* Simulator.AddObject (Simulator.fs:91-99): id = max id + 1 (SimulatorState.fs:90-92); `one` holds 1 body.
 * Every world owns a few spare slots on the device: while one is free the new body is a record write (tens of
 * microseconds, whatever the batch size); a full world lays the batch out again once.  Refusals (more than 1024 bodies,
 * a body outside the SpatialTree's space) leave the batch as it was. */
int ps_batch_add_body(void* handle, int32_t world, const ps_bodies_view* one);
/* Simulator.Reset (Simulator.fs:113-121): back to the state given at create; added bodies are dropped */
int ps_batch_reset(void* handle);

int ps_batch_stats(void* handle, ps_stats* out);
int ps_batch_num_bodies(void* handle, int32_t* n_bodies_total, int32_t* n_dynamic_total);
/* how this batch is scheduled: path 0 = fused per-world kernel, 1 = batch-wide stages;
 * lanes per world of the fused kernel; kernels launched by this handle so far (all entry points) */
int ps_batch_schedule(void* handle, int32_t* path, int32_t* lanes_per_world, int64_t* kernel_launches);
/* profiling aid: per world 8 x uint64 = SM cycles of the step phases (predict+cull spheres, near list, levels+sort,
 * wavefronts, rest), the slowest single step, sum of near-list lengths, sum of level counts */
int ps_batch_debug_cycles(void* handle, unsigned long long* out);

/* Broad phase on one large world (BASELINE config 5): sorted list of pairs (i<j) with >= 1 dynamic
 * body whose AABBs overlap strictly on all axes — SpatialTree.areIntersecting (SpatialTree.fs:39-41)
 * applied to SimulatorObject.getAABoundingBox (SimulatorObject.fs:58-71).
 * aabb_mode 0 = reference (sphere cube side = radius, Q5), 1 = conservative (half-extent = radius). */
int ps_broadphase_pairs(int32_t n_bodies, const int32_t* shape, const double* dims, const double* mass,
                        const double* pos, const double* rot, int32_t aabb_mode, int64_t* n_pairs,
                        int32_t* pair_ids /*2/pair*/, int64_t pair_capacity, double* kernel_ms);

#ifdef __cplusplus
}
#endif
#endif /* PS_B200_H */
