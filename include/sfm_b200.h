/*
 * sfm_b200 -- C ABI of the B200-native Social Force Model step.
 *
 * The reference (felixlutz/carla-social-force-model) is pure Python and has no FFI of its own: its seam is the Python
 * object protocol of forces.py / pedestrian_state.py / pedestrian_simulation.py.  This header is the boundary a
 * maintainer binds from that Python layer (ctypes); every entry point names the reference code it
 * replaces.  Plain pointers and sizes only.  All calls on one context must come from one host thread; all device work
 * is enqueued on the context's stream (sfm_set_stream), and calls that return host data synchronise that stream.
 *
 * Return value: 0 on success, non-zero on failure; sfm_last_error() returns the message of the calling thread's last
 * failure.  There is no CPU fallback: sfm_create fails when no sm_100 device is present.
 *
 * Host arrays are C-contiguous float64 unless stated otherwise, exactly the dtypes of the reference's PedState
 * (pedestrian_state.py:17-19).
 */
#ifndef SFM_B200_H
#define SFM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFM_ABI_VERSION 5

typedef struct sfm_ctx sfm_ctx;

/* Force classes in the order pedestrian_simulation.py:37-48 inserts them into its dict (= summation order, :81). */
enum sfm_force_class {
    SFM_FORCE_ACCELERATION = 0,     /* forces.py:35-53   AccelerationForce */
    SFM_FORCE_PEDESTRIAN = 1,       /* forces.py:56-117  PedestrianForce   */
    SFM_FORCE_BORDER = 2,           /* forces.py:120-179 BorderForce       */
    SFM_FORCE_STATIC_OBSTACLE = 3,  /* forces.py:182-283 ObstacleForce(dynamic=False) */
    SFM_FORCE_DYNAMIC_OBSTACLE = 4, /* forces.py:182-283 ObstacleForce(dynamic=True)  */
    SFM_FORCE_COUNT = 5
};

/* Pedestrian modes, ped_mode_manager.py:4-9.  Modes 2 and 3 switch the border force off (forces.py:176-177). */
enum sfm_ped_mode {
    SFM_IDLE = 0, SFM_WALKING_SIDEWALK = 1, SFM_CROSSING_ROAD = 2, SFM_ROAD_TO_SIDEWALK = 3, SFM_CHECKING_TRAFFIC = 4
};

/* One Moussaid parameter set: [pedestrian_force] forces.py:66-72, [static_/dynamic_obstacle_force] forces.py:196-206. */
typedef struct {
    double lambda_weight, A, gamma, n, n_prime, epsilon;
    double perception_threshold;          /* obstacle sets only (forces.py:206); ignored for pedestrians */
} sfm_moussaid_params;

/* Everything sfm_config.toml + the scenario's step_length configure on the hot path. */
typedef struct {
    double step_length;                   /* pedestrian_simulation.py:120 */
    double tau;                           /* forces.py:44  sfm_config['goal_force']['tau'], default 0.5 */
    double max_speed_factor;              /* pedestrian_state.py:15,72-73, default 1.3 */
    double border_a, border_b;            /* forces.py:135-136 */
    sfm_moussaid_params ped, static_obs, dynamic_obs;
    int32_t use_ped_radius;               /* forces.py:18 */
    int32_t enable[SFM_FORCE_COUNT];      /* [forces] switches, pedestrian_simulation.py:33-48 */
} sfm_params;

/* Device-time accounting of the kernels launched since the last sfm_reset_stats (CUDA events on the launch stream). */
typedef struct {
    int64_t launches;                     /* kernels of this library launched (all classes) */
    int64_t steps;                        /* fused steps executed */
    double ms_pairs;                      /* K1 all-pairs pedestrian force (incl. its partial-sum layout) */
    double ms_cells;                      /* K2a binning / cell-list builds */
    double ms_segments;                   /* K2b/K2c border + obstacle kernels */
    double ms_integrate;                  /* K3 acceleration + sum + clamp + Euler + restaging */
    int64_t pair_launches;                /* number of K1 launches inside ms_pairs */
    int64_t fixup_rows;                   /* rows the degenerate-pair repair path recomputed (0 for healthy crowds) */
    int64_t pair_evaluations;             /* pair terms K1 evaluated (padded slots included; one per unordered pair) */
    double ms_lifecycle;                  /* K4/K5/K6: mode machines, waypoint hand-over, vehicle rings, recorder */
    int64_t graph_replays;                /* ticks of sfm_step that ran as one CUDA graph launch (SFM_GRAPH=1 enables) */
    int64_t local_tile_pairs;             /* K1 tile pairs (256 x 256 rows) read through the origins of the partner tile's
                                             64-row runs: the one-subtraction "local" path, taken for spatially compact
                                             tiles (rows ordered along a space-filling curve); the rest took the
                                             double-single path.  Speed only -- both paths hold the force tolerance. */
} sfm_stats;

/* ---- lifetime ------------------------------------------------------------------------------------------------- */
int sfm_abi_version(void);
const char* sfm_last_error(void);
int sfm_device_count(int* count);
/* Creates a context on CUDA device `device`; fails unless its compute capability is 10.x. */
int sfm_create(int device, sfm_ctx** out);
int sfm_destroy(sfm_ctx* ctx);
/* `cuda_stream` is a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); NULL selects the context's own stream. */
int sfm_set_stream(sfm_ctx* ctx, void* cuda_stream);
int sfm_synchronize(sfm_ctx* ctx);

/* ---- configuration: replaces Force.__init__ (forces.py:14-18,41-44,62-72,127-136,189-206) ----------------------- */
int sfm_set_params(sfm_ctx* ctx, const sfm_params* params);
/* Origin subtracted (in float64) before positions are rounded to the float32 staging copies the pair kernel reads. */
int sfm_set_origin(sfm_ctx* ctx, double ox, double oy, double oz);
/* Row partition for multi-GPU runs: this context owns one of `world` equal blocks of `rows_pad` staged rows
 * (rows_pad a multiple of 256, >= the largest per-rank row count).  Default: world 1, rows_pad chosen by upload. */
int sfm_set_partition(sfm_ctx* ctx, int world, int rank, int64_t rows_pad);

/* ---- pedestrian state: replaces the PedState.state columns (pedestrian_state.py:17-19,45-77) -------------------- */
/* Uploads this context's rows.  loc, vel, next_waypoint: [n][3]; radius, target_speed: [n]; mode: uint8 [n]. */
int sfm_upload_state(sfm_ctx* ctx, int64_t n, const double* loc, const double* vel, const double* next_waypoint,
                     const double* radius, const double* target_speed, const uint8_t* mode);
/* Per-tick refresh of loc/vel from the simulator (PedState.update_state, pedestrian_state.py:79-81, batched). */
int sfm_update_kinematics(sfm_ctx* ctx, int64_t n, const double* loc, const double* vel);
/* Per-tick refresh of waypoint / target speed / mode (pedestrian_state.py:83-95).  NULL pointers leave a column as is. */
int sfm_update_targets(sfm_ctx* ctx, int64_t n, const double* next_waypoint, const double* target_speed,
                       const uint8_t* mode);
int sfm_download_state(sfm_ctx* ctx, int64_t n, double* loc, double* vel);

/* ---- point sets (CSR): replaces BorderForce.__init__ (forces.py:127-132) and ObstacleForce.update_obstacles /
 *      update_obstacle_velocities (forces.py:285-291).  offsets: int64 [count+1] into points [offsets[count]][2]. ---- */
int sfm_set_borders(sfm_ctx* ctx, int64_t n_sections, const double* section_center, const double* section_length,
                    const int64_t* offsets, const double* points);
/* which: SFM_FORCE_STATIC_OBSTACLE or SFM_FORCE_DYNAMIC_OBSTACLE.  velocities may be NULL (zeros, forces.py:212-213). */
int sfm_set_obstacles(sfm_ctx* ctx, int which, int64_t n_obstacles, const double* centers, const double* velocities,
                      const int64_t* offsets, const double* points);

/* ---- forces: replaces Force.get_force (forces.py:28-32) for each class --------------------------------------- */
/* Evaluates one force class on the current device state and copies it to out[n][3]. */
int sfm_force(sfm_ctx* ctx, int force_class, int64_t n, double* out);
/* Neighbour enumeration of a cutoff-limited class (border / static / dynamic): writes up to `capacity` triplets
 * (pedestrian row, section-or-obstacle index, nearest point index) in unspecified order and the total count. */
int sfm_enumerate_pairs(sfm_ctx* ctx, int force_class, int64_t capacity, int64_t* triplets, int64_t* count);
/* Work the reference does for a cutoff-limited class on the current state: `pairs` = (pedestrian, item) pairs inside the
 * cutoff (forces.py:149-151, :222-225), `point_evaluations` = the sum of those items' point counts, i.e. the distances
 * np.argmin ranges over (forces.py:154, :228).  The denominator of the distance-evaluations/s figure in bench.py. */
int sfm_count_point_evaluations(sfm_ctx* ctx, int force_class, int64_t* pairs, int64_t* point_evaluations);

/* ---- the fused tick: replaces PedestrianSimulation.tick's force sum and calculate_new_velocities
 *      (pedestrian_simulation.py:81-83,117-124; stateutils.py:18-23) plus, optionally, the position update the
 *      reference leaves to the CARLA server (run_simulation.py:77-87): x+ = x + dt * v+. --------------------------- */
/* With SFM_GRAPH=1 a single-rank tick's launches are captured once into a CUDA graph and replayed (re-captured whenever
 * an allocation, a size or a parameter changes; not while profiling or with fused routes).  Off by default: the tick is
 * bound by its chain of dependent kernels, not by launch overhead (profiles/small_n_steps_r1.log). */
int sfm_step(sfm_ctx* ctx, int n_steps, int integrate_positions);
/* Host-buffer tick: upload loc/vel, one step, download the new velocities (and positions when new_loc != NULL). */
int sfm_tick_host(sfm_ctx* ctx, int64_t n, const double* loc, const double* vel, double* new_vel, double* new_loc);
/* PedestrianSimulation.tick's device half on the reference's own pedestrian table (pedestrian_simulation.py:57-83): the
 * structured array `PedState.state` (pedestrian_state.py:17-19) is handed over as it lies in memory -- `records` points at
 * row 0, `stride` is the record size in bytes, `field_offsets` = byte offsets of loc, vel, next_waypoint, radius,
 * target_speed (4-byte aligned).  One H2D copy of the table, then on the device: unpack; with `tick_modes` the mode
 * bookkeeping of :63-73 (apply_current_mode, PedModeManager.tick, check_traffic -- needs sfm_set_mode_machines, and
 * sfm_set_traffic when vehicles exist); all enabled forces, their sum, the velocity update and clamp of :81-83,117-124.
 * The new velocities are written into the records' vel field (what `state[['id','vel']]` aliases, :123-124) and, with
 * `tick_modes`, the target speed the clamp used into target_speed (pedestrian_state.py:94-95).  counters4 (may be NULL)
 * receives sfm_lifecycle_counters after the tick, so the caller knows whether any machine changed mode.  Row count and
 * mode codes are those of the last sfm_upload_state / sfm_update_targets.
 * Identity column (ABI v4): the 8 bytes at `identity_offset` of every record -- the drop-in passes the `mode` column, whose
 * entries are the PedModeManager object pointers (pedestrian_state.py:17-19) -- tell whether the table still holds the
 * pedestrians the resident mode machines describe.  identity_mode 0: ignored; 1: adopted as the table's identity; 2:
 * compared with the adopted column on the device right after the upload -- on any difference (a spawn, a despawn, a
 * replaced object) *identity_changed = 1 and the call returns 0 with nothing else done: no mode step, no forces, the
 * records untouched; the caller rebuilds its mode table and calls again with identity_mode 1. */
int sfm_tick_records(sfm_ctx* ctx, int64_t n, void* records, int64_t stride, const int64_t* field_offsets,
                     double sim_time, int tick_modes, int64_t* counters4, int64_t identity_offset, int identity_mode,
                     int* identity_changed);
/* Page-lock a host range the caller owns (e.g. the PedState table) so that sfm_tick_records / sfm_tick_host copy from it
 * by DMA instead of through the driver's staging buffer; the caller keeps the memory alive until sfm_host_unregister. */
int sfm_host_register(void* ptr, size_t bytes);
int sfm_host_unregister(void* ptr);
/* Host-side helpers for AoS tables (no device work): copy one column out of / compare it with a packed array.  The
 * drop-in uses them on the `mode` column (object pointers) to notice that the table holds other objects than before. */
int sfm_host_column_gather(const void* records, int64_t stride, int64_t offset, int64_t width, int64_t n, void* packed);
int sfm_host_column_equal(const void* records, int64_t stride, int64_t offset, int64_t width, int64_t n,
                          const void* packed, int* equal);
/* calculate_new_velocities for a caller-composed force array force[n][3] (pedestrian_simulation.py:117-124 with
 * stateutils.cap_velocity, stateutils.py:18-23): v' = v + dt F clamped to target_speed * max_speed_factor, on the state
 * last uploaded; the new velocities replace the device's and are copied to new_vel[n][3]. */
int sfm_apply_force(sfm_ctx* ctx, int64_t n, const double* force, double* new_vel);
/* Total force / one class's force of the most recent step, [n][3]. */
int sfm_download_force(sfm_ctx* ctx, int64_t n, double* out);
int sfm_download_class_force(sfm_ctx* ctx, int force_class, int64_t n, double* out);

/* ---- multi-GPU plumbing: the staged planes (pos hi/lo, lambda*vel, radius, run-local pos, run/tile metadata) every
 *      rank all-gathers once per step ---
 * (SURVEY.md 8b sketched `sfm_comm_init(ctx, nccl_unique_id, rank, world)`: the library does not own a communicator --
 *  NCCL is driven by the caller (torch.distributed) over tensor views of the two buffers below, or bypassed altogether by
 *  the peer-memory exchange further down, where both collectives are fused into the library's own kernels.) */
/* Device pointer of the gather buffer [world][15][rows_pad] float32, and the size of one rank's block in bytes.
 * After each sfm_step the caller all-gathers block `rank` into every rank's buffer (NCCL, in place). */
int sfm_gather_buffer(sfm_ctx* ctx, void** device_ptr, size_t* bytes_per_rank);
/* Multi-GPU tick in two halves.  sfm_step_begin enqueues the pair accumulation (each unordered pedestrian pair is
 * evaluated once, by one rank, and contributes to both pedestrians' rows) and the cell-list forces; the caller then
 * reduce-scatters (integer sum) the accumulator -- block `rank` of [world][rows_pad][4] int64 stays on rank `rank` -- and
 * calls sfm_step_end (pair-force finish, K3), followed by the all-gather of the staged rows.  sfm_step == begin + end
 * on single-rank contexts. */
int sfm_step_begin(sfm_ctx* ctx);
int sfm_step_end(sfm_ctx* ctx, int integrate_positions);
/* Device pointer of the fixed-point force accumulator (int64 [world][rows_pad][4]) and the bytes of one rank block. */
int sfm_force_accumulator(sfm_ctx* ctx, void** device_ptr, size_t* bytes_per_rank);
/* Makes this rank's block of the gather buffer reflect the current master state (after an upload or refresh), so the
 * first all-gather can run before the first step. */
int sfm_stage(sfm_ctx* ctx);

/* ---- staged slot order.  Row order belongs to the caller: PedState.state is in spawn order (pedestrian_state.py:26-43,
 *      np.append) and every result of this library is per row.  Below the API the rows of a rank are STAGED for the pair
 *      kernel in the order of a Hilbert curve over their xy positions, so that runs of 64 consecutive staged slots are
 *      spatially compact and the kernel can read them through the run's own origin (one float32 subtraction per
 *      coordinate and pair instead of the three of the double-single form; sfm_stats.local_tile_pairs counts the tile
 *      pairs that did).  Speed only: neighbouring tiles and tiles with wide runs take the double-single path, both hold the force
 *      tolerance.  The float32 partial sums of the pair force follow the tile composition, so two contexts agree BIT FOR
 *      BIT only under the same order -- sfm_get_slot_order / sfm_set_slot_order carry it from one to the other (bench.py's
 *      single-GPU replay of a multi-rank tick); with the default interval 0 the order is the identity and nothing changes.
 *      No counterpart in the reference (its pair force materialises the dense (N, N-1) arrays of stateutils.py:32-75). */
/* Rebuild the order from the current positions every `ticks` ticks / restagings (0 = never: keep the current order). */
int sfm_set_reorder_interval(sfm_ctx* ctx, int ticks);
/* Rebuild the order now; the rows are restaged by the next step (multi-rank peer contexts: by the next sfm_stage). */
int sfm_reorder_slots(sfm_ctx* ctx);
/* slot_of_row: int32 [n], a permutation of [0, n) -- the staged slot of every row of this context. */
int sfm_get_slot_order(sfm_ctx* ctx, int64_t n, int32_t* slot_of_row);
int sfm_set_slot_order(sfm_ctx* ctx, int64_t n, const int32_t* slot_of_row);

/* ---- peer-memory exchange over NVLink (K7): the two collectives of the multi-GPU tick folded into the kernels next to
 *      them.  Set-up (once, after sfm_set_partition + sfm_upload_state on every rank): each rank exports three
 *      cudaIpcMemHandle_t (gather buffer, force accumulator, barrier flags: 3 x 64 bytes), the host side all-gathers
 *      them, each rank imports the table [world][3].  From then on sfm_stage and sfm_step_peer are COLLECTIVE calls:
 *      every rank of the box must make them in the same order. ------------------------------------------------------- */
int sfm_peer_export(sfm_ctx* ctx, void* handles);
int sfm_peer_import(sfm_ctx* ctx, const void* all_handles);
/* One tick = pair accumulation + cell-list forces; flag barrier; k1_sym_finish pulling every rank's partial accumulator
 * of its rows over NVLink (the reduce-scatter); K3 pushing the newly staged rows into every rank's gather buffer (the
 * all-gather); flag barrier.  No NCCL call on the path. */
int sfm_step_peer(sfm_ctx* ctx, int n_steps, int integrate_positions);
int sfm_peer_barrier(sfm_ctx* ctx);
/* barriers executed so far; timed_out != 0 when a barrier gave up waiting (~10 s, SFM_BARRIER_TIMEOUT_MS): bit r is set
 * for every rank r that did not arrive.  A timed-out exchange is not resumable -- results after it are invalid and later
 * barriers return immediately; destroy the contexts of all ranks (sfm_destroy) and create new ones. */
int sfm_peer_status(sfm_ctx* ctx, int64_t* barriers, int* timed_out);

/* ---- lifecycle on the device: what PedestrianSimulation.tick and SimulationRunner.tick do in
 *      interpreter loops around the forces.  All per-row tables are dropped by the next sfm_upload_state. ------------- */
/* f1.  The persistent fields of every pedestrian's PedModeManager (ped_mode_manager.py:18-28): initial_target_speed,
 * crossing_speed (= crossing_speed_factor * target_speed), crossing_safety_margin, target_speed (of the mode, :25),
 * next_mode_time; current_mode is the `mode` column of sfm_upload_state.  waiting_time is :28 (5 s). */
int sfm_set_mode_machines(sfm_ctx* ctx, int64_t n, const double* initial_target_speed, const double* crossing_speed,
                          const double* crossing_safety_margin, const double* mode_target_speed,
                          const double* next_mode_time, double waiting_time);
/* Vehicle kinematics for the gap-acceptance test when the rings come from the host (pedestrian_simulation.py:108-113:
 * obstacle positions, velocities, extents); sfm_set_vehicles supplies the same data itself.  extents: [V][2]. */
int sfm_set_traffic(sfm_ctx* ctx, int64_t n_vehicles, const double* centers, const double* velocities,
                    const double* extents);
/* One tick of every machine: PedState.apply_current_mode (pedestrian_state.py:94-95: target_speed <- mode.target_speed),
 * PedModeManager.tick (ped_mode_manager.py:30-35) and the gap acceptance of every CHECKING_TRAFFIC pedestrian against
 * every vehicle (pedestrian_simulation.py:67-73, check_traffic.py:7-61) -> set_mode(CROSSING_ROAD). */
int sfm_tick_modes(sfm_ctx* ctx, double sim_time);
/* Any pointer may be NULL.  target_speed is the PedState column (what the speed clamp reads). */
int sfm_download_modes(sfm_ctx* ctx, int64_t n, uint8_t* mode, double* mode_target_speed, double* next_mode_time,
                       double* target_speed);
/* f3.  Remaining waypoints per pedestrian (SimulationRunner.waypoint_dict, run_simulation.py:118-125) as CSR:
 * offsets int64 [n+1] into waypoints [W][3] and crossing uint8 [W] (the (waypoint, crossing_road) tuples of
 * pedestrian_state.py:83-92).  With `fused` != 0 every sfm_step / sfm_step_end performs the arrival test
 * (pedestrian_simulation.py:88-97, at the positions the forces were evaluated at) and the hand-over inside K3. */
int sfm_set_routes(sfm_ctx* ctx, int64_t n, const int64_t* offsets, const double* waypoints, const uint8_t* crossing,
                   double distance_threshold, int fused);
/* The same arrival test + hand-over as a call of its own, at the current device positions. */
int sfm_advance_waypoints(sfm_ctx* ctx);
/* cursor: entries of each pedestrian's route already handed out; finished: arrived with nothing left (run_simulation.py:127). */
int sfm_download_routes(sfm_ctx* ctx, int64_t n, int64_t* cursor, uint8_t* finished, double* next_waypoint);
/* Appends m pedestrians to the crowd: PedestrianSimulation.spawn_pedestrian / PedState.add_pedestrian
 * (pedestrian_simulation.py:99-100, pedestrian_state.py:26-40), batched.  The first six arrays are those of
 * sfm_upload_state; the five machine arrays (sfm_set_mode_machines) are required when the context carries mode machines,
 * the route CSR (offsets int64 [m+1], waypoints, crossing; sfm_set_routes) when it carries routes.  Existing rows keep
 * their indices.  Synchronises the stream.  Single-rank contexts only. */
int sfm_append_pedestrians(sfm_ctx* ctx, int64_t m, const double* loc, const double* vel, const double* next_waypoint,
                           const double* radius, const double* target_speed, const uint8_t* mode,
                           const double* initial_target_speed, const double* crossing_speed,
                           const double* crossing_safety_margin, const double* mode_target_speed,
                           const double* next_mode_time, const int64_t* route_offsets, const double* waypoints,
                           const uint8_t* crossing);
/* Removes every pedestrian whose `finished` flag is set (run_simulation.py:127-132 with despawn_on_arrival; row order
 * preserved like PedState.remove_pedestrian, pedestrian_state.py:42-43) from all per-row tables and restages.  Returns
 * the new row count; later uploads / downloads use it.  Synchronises the stream.  Single-rank contexts only. */
int sfm_despawn_finished(sfm_ctx* ctx, int64_t* n_after, int64_t* n_removed);
/* out4: crossings started, idle wake-ups, waypoint hand-overs, pedestrians finished -- since the context was created. */
int sfm_lifecycle_counters(sfm_ctx* ctx, int64_t* out4);
/* f2.  Dynamic-obstacle set generated on the device: obstacles.py:297-329 (get_dynamic_obstacles) with the ellipse of
 * :269-281 -- max(6, int((2 ex + 2 ey) / resolution)) points per vehicle on semi-axes extent * size_factor, rotated by
 * yaw (degrees) about the centre.  Replaces sfm_set_obstacles(SFM_FORCE_DYNAMIC_OBSTACLE, ...) + sfm_set_traffic. */
int sfm_set_vehicles(sfm_ctx* ctx, int64_t n_vehicles, const double* centers, const double* yaw_deg,
                     const double* velocities, const double* extents, double resolution, double size_factor);
/* Headless stand-in for the simulator's vehicle actors: centres += velocity * dt, rings regenerated, cells rebuilt. */
int sfm_advance_vehicles(sfm_ctx* ctx, double dt);
/* centers [V][2], offsets int64 [V+1], points [point_capacity][2]; any pointer may be NULL. */
int sfm_download_vehicles(sfm_ctx* ctx, int64_t n_vehicles, double* centers, int64_t* offsets, int64_t point_capacity,
                          double* points);
/* f4.  Device-resident recorder: PedState.record_current_state (pedestrian_state.py:100-104) keeps the columns
 * pedestrian.csv is written from (output_generator.py:35-52): per frame [n][4] float64 (x, y, v_x, v_y) + uint8 mode. */
int sfm_record_begin(sfm_ctx* ctx, int64_t capacity_frames);
int sfm_record_frame(sfm_ctx* ctx, double sim_time);
int sfm_download_frames(sfm_ctx* ctx, int64_t first, int64_t count, double* xyv, uint8_t* mode, double* times,
                        int64_t* frames_recorded);

/* ---- accounting -------------------------------------------------------------------------------------------------- */
int sfm_set_profiling(sfm_ctx* ctx, int enabled);
int sfm_reset_stats(sfm_ctx* ctx);
int sfm_get_stats(sfm_ctx* ctx, sfm_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* SFM_B200_H */
