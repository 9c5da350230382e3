"""CSV writer with the call surface of the reference's ``output_generator.py``.

Same constructor, same four methods, same file names, headers and column order (output_generator.py:32-110); rows are
produced column-wise (``ndarray.tolist()`` + ``csv.writer.writerows``) instead of one interpreter iteration per field,
and the pedestrian frames may come from the device-resident recorder (``sfm_record_frame``) instead of the per-tick
structured-array snapshots (pedestrian_state.py:100-104).  Number formatting is ``str(float)``, which is what
``csv.writer`` applies to the numpy scalars the reference hands it.
"""
import csv
import os
import time

import numpy as np


def _mode_column(modes):
    return [str(int(getattr(m, 'current_mode', m))) for m in modes]


class OutputGenerator:
    def __init__(self, ped_sim, output_path, scenario_name):
        self.scene = ped_sim
        self.ped_states = self.scene.peds.all_states
        self.veh_states = self.scene.all_dyn_obs_states
        self.static_obstacles = self.scene.static_obstacles
        self.borders = self.scene.borders
        self.output_path = output_path
        time_stamp = time.strftime('%Y%m%d-%H%M%S')
        dir_name = time_stamp + '-' + scenario_name if scenario_name else time_stamp
        self.output_dir = os.path.join(output_path, dir_name)
        os.makedirs(self.output_dir, exist_ok=True)

    def _open(self, name):
        return open(os.path.join(self.output_dir, name), 'w', encoding='UTF8', newline='')

    def generate_ped_csv(self, device_frames=None, ped_ids=None):
        """pedestrian.csv: ped_id,frame,time,x,y,v_x,v_y,mode  (output_generator.py:32-52).

        ``device_frames`` = (times [F], xyv [F, n, 4], mode [F, n]) as ``Context.download_frames`` returns them, with
        ``ped_ids`` [n] the integer suffixes of the pedestrian names -- or a list of (times, xyv, mode, ids) segments when
        the crowd changed during the run (``HeadlessRunner.recorded_frames``: one segment per row set, frame numbers run
        on); default: the host snapshots in ``peds.all_states``.
        """
        with self._open('pedestrian.csv') as f:
            writer = csv.writer(f)
            writer.writerow(['ped_id', 'frame', 'time', 'x', 'y', 'v_x', 'v_y', 'mode'])
            if device_frames is not None:
                segments = device_frames if isinstance(device_frames, list) else [tuple(device_frames) + (ped_ids,)]
                frame = 0
                for times, xyv, mode, seg_ids in segments:
                    ids = np.asarray(seg_ids).tolist()
                    for k, sim_time in enumerate(np.asarray(times).tolist()):
                        cols = xyv[k]
                        writer.writerows(zip(ids, [frame] * len(ids), [sim_time] * len(ids), cols[:, 0].tolist(),
                                             cols[:, 1].tolist(), cols[:, 2].tolist(), cols[:, 3].tolist(),
                                             mode[k].tolist()))
                        frame += 1
                return
            for frame, (sim_time, state) in enumerate(self.ped_states.items()):
                n = len(state)
                ids = [int(name.split('_')[-1]) for name in state['name'].tolist()]
                loc, vel = state['loc'], state['vel']
                writer.writerows(zip(ids, [frame] * n, [sim_time] * n, loc[:, 0].tolist(), loc[:, 1].tolist(),
                                     vel[:, 0].tolist(), vel[:, 1].tolist(), _mode_column(state['mode'])))

    def generate_veh_csv(self):
        """vehicle.csv: veh_id,frame,time,x,y,heading,vel,ext_x,ext_y  (output_generator.py:54-75)."""
        with self._open('vehicle.csv') as f:
            writer = csv.writer(f)
            writer.writerow(['veh_id', 'frame', 'time', 'x', 'y', 'heading', 'vel', 'ext_x', 'ext_y'])
            for frame, (sim_time, state) in enumerate(self.veh_states.items()):
                n = len(state)
                heading = np.deg2rad(state['heading'])
                speed = np.linalg.norm(state['vel'], axis=-1)
                writer.writerows(zip(state['id'].tolist(), [frame] * n, [sim_time] * n, state['loc'][:, 0].tolist(),
                                     state['loc'][:, 1].tolist(), heading.tolist(), speed.tolist(),
                                     state['extent'][:, 0].tolist(), state['extent'][:, 1].tolist()))

    def generate_borders_csv(self):
        """borders.csv: x,y  (output_generator.py:77-91)."""
        with self._open('borders.csv') as f:
            writer = csv.writer(f)
            writer.writerow(['x', 'y'])
            for border in self.borders:
                border = np.asarray(border, dtype=np.float64)
                writer.writerows(zip(border[:, 0].tolist(), border[:, 1].tolist()))

    def generate_obstacles_csv(self):
        """obstacles.csv: obs_id,obs_pos_x,obs_pos_y,x,y  (output_generator.py:93-110)."""
        with self._open('obstacles.csv') as f:
            writer = csv.writer(f)
            writer.writerow(['obs_id', 'obs_pos_x', 'obs_pos_y', 'x', 'y'])
            for obs_id, (pos, border) in enumerate(self.static_obstacles):
                border = np.asarray(border, dtype=np.float64)
                n = len(border)
                writer.writerows(zip([obs_id] * n, [float(pos[0])] * n, [float(pos[1])] * n, border[:, 0].tolist(),
                                     border[:, 1].tolist()))
