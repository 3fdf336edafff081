/* ps_snapshot.h — golden-trajectory snapshot format "PSTJ", version 1.
 *
 * The reference has no on-disk or wire format at all: `SimulatorState` (PhysicsSimulatorCore/SimulatorState.fs:11-19)
 * only lives in memory.  To exchange trajectories between the three implementations of the step path — the F#
 * reference on a box with .NET, the CPU oracle (oracle/ps_oracle.cpp) and the CUDA path behind include/ps_b200.h —
 * one file holds everything the parity bar compares: the configuration POD, the constant body parameters, and per
 * recorded step the SoA state, the ordered candidate list, the colliding-pair set, the contact lists and the
 * contact-feature ids.
 *
 * Writers/readers: physicssimulator_b200/snapshot.py (Python; used by tests/ and tools/compare_snapshots.py) and
 * the F# module in INTEGRATION.md §6 (for the headless .NET driver of SURVEY.md §8(d) plan A).  Everything is
 * little-endian, no padding beyond what is declared, arrays in the layout of include/ps_b200.h (body index fastest
 * inside a world, worlds concatenated, vectors xyz-interleaved, orientation row-major).
 *
 *   file   := ps_snapshot_header
 *             int32   world_offset[n_worlds + 1]
 *             constants                                  (n = world_offset[n_worlds] bodies)
 *             frame * n_frames
 *   constants := int32 shape[n]  double dims[3n]  double mass[n]  double inertia_inv[3n]  double force[3n]
 *                double torque[3n]  double elasticity[n]  double mu_kinetic[n]  double mu_static[n]
 *   frame  := ps_snapshot_frame_header
 *             double pos[3n]  double vel[3n]  double rot[9n]  double ang_mom[3n]
 *             per world, when PS_SNAP_HAS_CANDIDATES:  int32 n_cand,  int32 cand[2 * n_cand]      (ordered)
 *             per world, when PS_SNAP_HAS_CONTACTS:    int32 n_pairs, int32 n_contacts,
 *                                                      int32 pair_ids[2 * n_pairs]   (i < j, processing order)
 *                                                      int32 first[n_pairs + 1]      (prefix into the contact arrays)
 *                                                      uint32 feature[n_contacts]    (ps_b200.h, contact-feature id)
 *                                                      double normal[3 * n_contacts] double position[3 * n_contacts]
 *                                                      double penetration[n_contacts]
 * `mass` = +inf marks a static body (Entities/Base.fs:13-20).  The candidate/contact sections describe the step that
 * PRODUCED the frame's state (the `Collisions` map of SimulatorState.fs:17 after that step); frame 0 (step 0, the
 * initial state) carries empty sections.
 */
#ifndef PS_SNAPSHOT_H
#define PS_SNAPSHOT_H

#include <stdint.h>
#include "ps_b200.h"

#define PS_SNAP_MAGIC 0x4a545350u       /* "PSTJ" */
#define PS_SNAP_FRAME_MAGIC 0x4d415246u /* "FRAM" */
#define PS_SNAP_VERSION 1u

enum {
    PS_SNAP_HAS_CANDIDATES = 1u, /* ordered candidate list per world and frame */
    PS_SNAP_HAS_CONTACTS = 2u    /* colliding pairs, contact lists, feature ids per world and frame */
};

/* who produced the file: the comparison tool prints it and picks the tolerance policy from it (DESIGN.md §2:
 * oracle <-> CUDA is bit-exact; either of them <-> the .NET reference is a stated FP64 tolerance on states and
 * equality on the sets as long as no discrete decision has diverged) */
enum {
    PS_SNAP_PRODUCER_UNKNOWN = 0,
    PS_SNAP_PRODUCER_REFERENCE_DOTNET = 1, /* PhysicsSimulatorCore itself, through the driver of INTEGRATION.md §6 */
    PS_SNAP_PRODUCER_ORACLE = 2,           /* oracle/ps_oracle.cpp */
    PS_SNAP_PRODUCER_CUDA = 3              /* libps_b200.so */
};

#pragma pack(push, 1)
typedef struct {
    uint32_t magic;        /* PS_SNAP_MAGIC */
    uint32_t version;      /* PS_SNAP_VERSION */
    uint32_t header_bytes; /* sizeof(ps_snapshot_header): readers skip what they do not know */
    uint32_t flags;        /* PS_SNAP_HAS_* */
    uint32_t producer;     /* PS_SNAP_PRODUCER_* */
    uint32_t n_worlds;
    uint32_t n_bodies;     /* == world_offset[n_worlds] */
    uint32_t n_frames;
    uint32_t config_bytes; /* sizeof(ps_step_config) at the time of writing */
    uint32_t reserved;
    ps_step_config config; /* the configuration the trajectory was stepped with */
    char scene[64];        /* free text, NUL padded: scene name / seed */
} ps_snapshot_header;

typedef struct {
    uint32_t magic; /* PS_SNAP_FRAME_MAGIC */
    int32_t step;   /* number of steps taken from the initial state */
} ps_snapshot_frame_header;
#pragma pack(pop)

#endif /* PS_SNAPSHOT_H */
