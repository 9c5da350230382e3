"""Device-resident run loop: what ``SimulationRunner.tick`` does every step (run_simulation.py:59-132), CARLA stubbed.

Per tick, all on the device behind the C ABI:

1. vehicles -- either the host hands over the 6-tuple of pedestrian_simulation.py:108-115 (rings as CARLA would produce
   them) or the library moves the centres ballistically and regenerates the ellipse rings itself (``sfm_set_vehicles`` /
   ``sfm_advance_vehicles``; obstacles.py:269-281,297-329);
2. ``sfm_tick_modes`` -- apply_current_mode, the mode machines' tick and the gap acceptance (pedestrian_simulation.py:63-73);
3. ``sfm_step`` -- the five forces, their sum, the velocity update with the speed clamp, the arrival test + waypoint
   hand-over at the positions the forces saw (run_simulation.py:118-125), then the stub integrator x += dt v;
4. optionally ``sfm_record_frame`` every ``record_every`` ticks (pedestrian_state.py:100-104).

Nothing here computes: the class only sequences C-ABI calls.
"""
from __future__ import annotations

import numpy as np

from . import native

IDLE = 0


class HeadlessRunner:
    def __init__(self, sfm_config, workload, life, device=0, device_vehicles=False, record_every=0, record_capacity=0,
                 ctx=None, reorder_every=32):
        w = workload
        self.w, self.life = w, life
        self.dt = w.step_length
        self.ctx = ctx or native.Context(device)
        c = self.ctx
        c.set_params(native.params_from_config(sfm_config, w.step_length))
        c.set_reorder_interval(reorder_every)              # staged slot order (csrc/k8_order.cuh): speed only
        spawn_tick = getattr(life, 'spawn_tick', None)
        self.spawn_tick = np.zeros(w.n, dtype=np.int64) if spawn_tick is None else np.asarray(spawn_tick)
        first = np.nonzero(self.spawn_tick == 0)[0]
        if len(first) < w.n:                               # late spawners join through sfm_append_pedestrians
            import copy
            full, w = w, copy.copy(w)
            for name in ('loc', 'vel', 'next_waypoint', 'radius', 'target_speed', 'mode'):
                setattr(w, name, getattr(full, name)[first])
            life = copy.copy(life)
            life.routes = [self.life.routes[i] for i in first]
            for name in ('crossing_speed_factor', 'crossing_safety_margin', 'idle'):
                setattr(life, name, np.asarray(getattr(self.life, name))[first])
        c.upload_state(w.loc, w.vel, w.next_waypoint, w.radius, w.target_speed, w.mode)
        if len(w.borders):
            c.set_borders(w.borders, w.section_center, w.section_length)
        if len(w.static_obstacles):
            c.set_obstacles(native.STATIC_OBSTACLE, [p for p, _ in w.static_obstacles], [r for _, r in w.static_obstacles])
        # PedModeManager.__init__ (ped_mode_manager.py:18-28) for every pedestrian, then set_mode(IDLE) at t = 0 for the
        # idle ones: target speed 0, wake-up at waiting_time
        crossing_speed = np.asarray(life.crossing_speed_factor, dtype=np.float64) * w.target_speed
        mode_speed = np.where(life.idle, 0.0, w.target_speed)
        next_time = np.where(life.idle, 0.0 + 5.0, -1.0)
        mode = np.where(life.idle, IDLE, w.mode).astype(np.uint8)
        c.update_targets(mode=mode)
        c.set_mode_machines(w.target_speed, crossing_speed, life.crossing_safety_margin, mode_speed, next_time, 5.0)
        c.set_routes(life.routes, life.waypoint_threshold, fused=True)
        self.device_vehicles = device_vehicles
        self.has_vehicles = w.veh_center is not None and len(w.veh_center) > 0
        if self.has_vehicles and device_vehicles:
            c.set_vehicles(w.veh_center, w.veh_yaw, w.veh_vel, w.veh_extent, w.veh_resolution)
        self.step_index = 0
        self.ids = first.copy()                            # original row of every pedestrian in the crowd
        self._finished_seen = 0
        self.record_every = record_every
        self.record_capacity = record_capacity
        self.all_dyn_obs_states = {}
        self._segments = []                                # recorded frames of earlier row sets: (times, xyv, mode, ids)
        if record_every:
            c.record_begin(record_capacity)

    def _flush_recorder(self):
        """The device recorder holds frames of ONE row count.  Before the crowd changes (spawn / despawn) the frames
        recorded so far move to the host together with the ids they belong to (pedestrian_state.py:100-104 keeps a full
        snapshot per tick, so the reference records across spawns and despawns); ``_rearm_recorder`` follows the change."""
        if not self.record_every:
            return
        times, xyv, mode = self.ctx.download_frames()
        if len(times):
            self._segments.append((times, xyv, mode, self.ids.copy()))

    def _rearm_recorder(self):
        if self.record_every:
            self.ctx.record_begin(self.record_capacity)

    def recorded_frames(self):
        """Every frame recorded so far as a list of (times [F], xyv [F, n, 4], mode [F, n], ids [n]) segments."""
        out = list(self._segments)
        if self.record_every:
            times, xyv, mode = self.ctx.download_frames()
            if len(times):
                out.append((times, xyv, mode, self.ids.copy()))
        return out

    def _spawn(self, rows):
        """PedSpawner.tick -> PedestrianSimulation.spawn_pedestrian (pedestrian_spawner.py:238-241), batched."""
        w, life = self.w, self.life
        speed = w.target_speed[rows]
        idle = np.asarray(life.idle)[rows]
        machines = (speed, np.asarray(life.crossing_speed_factor)[rows] * speed, np.asarray(life.crossing_safety_margin)[rows],
                    np.where(idle, 0.0, speed), np.where(idle, self.step_index * self.dt + 5.0, -1.0))
        self.ctx.append_pedestrians(w.loc[rows], w.vel[rows], w.next_waypoint[rows], w.radius[rows], speed,
                                    np.where(idle, IDLE, w.mode[rows]).astype(np.uint8), machines,
                                    [life.routes[i] for i in rows])
        self.ids = np.concatenate((self.ids, rows))

    def tick(self):
        c, k = self.ctx, self.step_index
        if k > 0:
            late = np.nonzero(self.spawn_tick == k)[0]
            if len(late):
                self._flush_recorder()
                self._spawn(late)
                self._rearm_recorder()
        if self.has_vehicles:
            if self.device_vehicles:
                if k > 0:
                    c.advance_vehicles(self.dt)
            else:
                veh = self.w.vehicles_at(k)
                c.set_obstacles(native.DYNAMIC_OBSTACLE, veh[1], veh[5], veh[3])
                c.set_traffic(veh[1], veh[3], veh[4])
        sim_time = k * self.dt
        c.tick_modes(sim_time)
        if self.record_every and k % self.record_every == 0:
            c.record_frame(sim_time)                      # after the machines ticked, before the forces (:75)
            if self.has_vehicles:                         # record_dyn_obstacle_states (pedestrian_simulation.py:129-140)
                centres = c.download_vehicles()[0] if self.device_vehicles else np.asarray(self.w.vehicles_at(k)[1])
                st = np.zeros(len(centres), dtype=[('id', 'i4'), ('loc', 'f8', (2,)), ('heading', 'f8'),
                                                    ('vel', 'f8', (2,)), ('extent', 'f8', (2,))])
                st['id'] = np.arange(1000, 1000 + len(centres))
                st['loc'], st['heading'], st['vel'], st['extent'] = centres, self.w.veh_yaw, self.w.veh_vel, self.w.veh_extent
                self.all_dyn_obs_states[sim_time] = st
        c.step(1, integrate_positions=True)
        if self.life.despawn_on_arrival and c.lifecycle_counters()['finished'] > self._finished_seen:
            # run_simulation.py:127-132.  The counter read-back (32 bytes) is the only per-tick synchronisation; the mask
            # is fetched only on ticks on which somebody actually finished.
            _, finished, _ = c.download_routes()
            self._flush_recorder()
            self.ids = self.ids[~finished]
            self._finished_seen += c.despawn_finished()
            self._rearm_recorder()
        self.step_index += 1

    def run(self, n_steps):
        for _ in range(n_steps):
            self.tick()

    def write_csv(self, output_path, scenario_name):
        """The reference's four result files (output_generator.py) from the device-recorded frames of this run."""
        import types
        from output_generator import OutputGenerator
        sim = types.SimpleNamespace(peds=types.SimpleNamespace(all_states={}), all_dyn_obs_states=self.all_dyn_obs_states,
                                    static_obstacles=list(self.w.static_obstacles), borders=list(self.w.borders))
        gen = OutputGenerator(sim, output_path, scenario_name)
        gen.generate_ped_csv(device_frames=self.recorded_frames())
        gen.generate_veh_csv()
        gen.generate_borders_csv()
        gen.generate_obstacles_csv()
        return gen.output_dir

    def snapshot(self):
        loc, vel = self.ctx.download_state()
        modes = self.ctx.download_modes()
        cursor, finished, wp = self.ctx.download_routes()
        return dict(loc=loc, vel=vel, wp=wp, cursor=cursor, finished=finished, **modes)
