"""VideoMaker (reference real_robots/videomaker.py:11-190) on the batched rasteriser.

The reference films the scene with pybullet's OpenGL renderer through an `EnvCamera` (env.py:470-513) posed at
distance 1.0, yaw 90, pitch -45, target [-0.3, 0, 0.4], fov 90 (videomaker.py:25-26), every `frame_freq` = 8 simulation
steps (25 fps of 200 Hz), pastes the GOAL / START retina insets, optionally draws step / score text, and writes XVID
.avi files named like the reference's.  Here the frame comes from rr_render_camera (same camera matrices, TinyRenderer
conventions instead of OpenGL shading — documented deviation).  `interval` (pyinterval) is not a dependency: any object
supporting `in` and `len` is accepted, and `Frames` below stands in for it.
"""
import time

import numpy as np

from . import cameras

VIDEO_WIDTH = 320
VIDEO_HEIGHT = 240


class Frames:
    """Stand-in for pyinterval's `interval`: Frames([a, b], [c, d], 7) contains a<=x<=b, c<=x<=d and x == 7."""

    def __init__(self, *parts):
        self.parts = [(float(p[0]), float(p[1])) if isinstance(p, (list, tuple)) else (float(p), float(p)) for p in parts]

    def __contains__(self, x):
        return any(lo <= x <= hi for lo, hi in self.parts)

    def __len__(self):
        return len(self.parts)


class EnvCamera:
    """env.py:470-513: yaw / pitch / roll camera about a target point; render(env) -> u8[240, 320, 3]."""

    def __init__(self, distance, yaw, pitch, roll, pos, fov=80, width=320, height=240):
        self.dist, self.yaw, self.pitch, self.roll, self.pos = distance, yaw, pitch, roll, pos
        self.fov, self.render_width, self.render_height = fov, width, height

    def matrices(self):
        view = cameras.view_from_yaw_pitch_roll(self.pos, self.dist, self.yaw, self.pitch, self.roll, 2)
        proj = cameras.projection_fov(self.fov, float(self.render_width) / self.render_height, 0.1, 100.0)
        return view, proj

    def render(self, env):
        if (self.render_width, self.render_height) != (VIDEO_WIDTH, VIDEO_HEIGHT):
            raise ValueError("the rasteriser is built for 320x240 images")
        sim = getattr(env, "_sim", None) or getattr(env, "sim", None) or env
        view, proj = self.matrices()
        rgb, _, _ = sim.render_camera(view, proj)
        return rgb[0].cpu().numpy()


class VideoMaker:
    def __init__(self, env, intrinsic=None, extrinsic=None, debug=False):
        self.env = env
        self.camera = EnvCamera(1.0, 90, -45, 0, [-0.3, 0, .4], fov=90, width=VIDEO_WIDTH, height=VIDEO_HEIGHT)
        self.seed = np.random.randint(100000)
        self.video = None
        self.video_fps = 25
        self.speed_up = 1
        self.frame_freq = int((200.0 / self.video_fps) * self.speed_up)
        self.debug = debug
        self.trial_number = None
        self.files = []
        self.intrinsic_frames = self._select(intrinsic, self.get_intrinsic_frames, "intrinsic")
        self.extrinsic_trials = self._select(extrinsic, self.get_extrinsic_trials, "extrinsic")

    @staticmethod
    def _select(arg, default, what):
        if not arg:
            return Frames()
        if isinstance(arg, bool):
            return default()
        if hasattr(arg, "__contains__") and hasattr(arg, "__len__"):
            return arg
        raise Exception("VideoMaker %s param has to be either None/False, an interval or True" % what)

    def get_intrinsic_frames(self):                     # videomaker.py:60-65: first, middle and last minute of video
        int_steps = self.env.intrinsic_timesteps
        one_min_frames = 60 * self.video_fps * self.frame_freq
        return Frames([0, one_min_frames], [int_steps / 2, int_steps / 2 + one_min_frames], [int_steps - one_min_frames, int_steps])

    def get_extrinsic_trials(self):                     # videomaker.py:67-74: five random trials
        ext_trials = int(self.env.extrinsic_trials)
        if ext_trials <= 0:
            return Frames()
        return Frames(*[int(t) for t in np.random.choice(ext_trials, min(ext_trials, 5), replace=False)])

    # ------------------------------------------------------------------ file handling
    def _open(self, suffix):
        import cv2
        t = time.strftime("%Y,%m,%d,%H,%M").split(',')
        name = "Simulation-{}-y{}-m{}-d{}-h{}-m{}-{}.avi".format(self.seed, *t, suffix)
        self.video = cv2.VideoWriter(name, cv2.VideoWriter_fourcc(*'XVID'), self.video_fps, (VIDEO_WIDTH, VIDEO_HEIGHT), isColor=True)
        self.files.append(name)

    def _write(self, image):
        import cv2
        frame = np.asarray(image, dtype=np.uint8).reshape(VIDEO_HEIGHT, VIDEO_WIDTH, 3)
        self.video.write(cv2.cvtColor(frame, cv2.COLOR_RGB2BGR))

    def _close(self):
        if self.video is not None:
            self.video.release()
            self.video = None

    # ------------------------------------------------------------------ intrinsic phase
    def start_intrinsic(self):
        if len(self.intrinsic_frames) > 0:
            self._open("intrinsic")

    def update_intrinsic(self, steps):
        if steps in self.intrinsic_frames and steps % self.frame_freq == 0 and self.video is not None:
            from PIL import Image
            camera = Image.fromarray(self.camera.render(self.env))
            if self.debug:
                self.addDebugInfo(camera, steps, None, False)
            self._write(camera)

    def end_intrinsic(self):
        if len(self.intrinsic_frames) > 0:
            self._close()

    # ------------------------------------------------------------------ extrinsic phase
    def makeInset(self, image, text, right):
        from PIL import Image, ImageDraw
        img = Image.fromarray(np.asarray(image, dtype=np.uint8))
        i_width, i_height = int(VIDEO_WIDTH / 3), int(VIDEO_HEIGHT / 3)
        img = img.resize((i_width, i_height))
        d = ImageDraw.Draw(img)
        box = d.textbbox((0, 0), text)
        w, h = box[2] - box[0], box[3] - box[1]
        d.text((int((i_width - w) / 2), int((i_height * 0.75) - h / 2)), text, fill=(0, 0, 0))
        return img

    def start_trial(self, observation, trial_number):
        self.trial_number = trial_number
        if self.trial_number in self.extrinsic_trials:
            self._open("trial-{}".format(self.trial_number))
            self.goal = self.makeInset(observation['goal'], "GOAL", True)
            self.start = self.makeInset(observation['retina'], "START", False)

    def extrinsic_trial(self, observation, action, steps, score_object):
        if self.trial_number in self.extrinsic_trials and steps % self.frame_freq == 0 and self.video is not None:
            from PIL import Image
            camera = Image.fromarray(self.camera.render(self.env))
            camera.paste(self.goal, (VIDEO_WIDTH - int(VIDEO_WIDTH / 3), 0))
            camera.paste(self.start, (0, 0))
            if self.debug:
                self.addDebugInfo(camera, steps, score_object, True)
            self._write(camera)

    def end_trial(self):
        if self.trial_number in self.extrinsic_trials:
            self._close()

    def addDebugInfo(self, camera, steps, score_object, extrinsic):
        from PIL import ImageDraw
        d = ImageDraw.Draw(camera)
        if extrinsic:
            h = int(VIDEO_HEIGHT / 3) + 3
            w = VIDEO_WIDTH - int(VIDEO_WIDTH / 3) + 3
            d.text((3, h), "Trial: " + str(self.trial_number) + "\nStep: " + str(steps), fill=(0, 0, 0))
            if self.trial_number and score_object:
                d.text((w, h), "Score: " + str(score_object["score_total"])[:5] + "\nScore 2D: " + str(score_object['score_2D'])[:5] +
                       "\nScore 2.5D: " + str(score_object['score_2.5D'])[:5] + "\nScore 3D: " + str(score_object['score_3D'])[:5], fill=(0, 0, 0))
        else:
            d.text((3, 3), "Intrinsic phase" + "\nStep: " + str(steps), fill=(0, 0, 0))
