"""Python mirror of the reference's MatchGPULib interface (src/gpu_matcher/MatchGPULib.h:6-47) over the
C-ABI: same method names, argument meaning and result layouts, numpy arrays instead of malloc'd planes."""
import numpy as np

from . import capi


class MatchGPULib:
    def __init__(self, argv=None):
        argv = list(argv or [])
        self.foveatedmatching = 0  # MatchGPULib.cpp:255-257
        self.fovH = 0
        self.fovW = 0
        self.foveatelevel = int(argv[2]) if len(argv) > 2 else 7  # MatchGPULib.cpp:259-264
        device = -1
        for a in argv[1:]:
            if isinstance(a, str) and a.startswith("-device="):
                device = int(a[8:])
        self._h = capi.Handle(fovea_level=self.foveatelevel, device=device)

    def close(self):
        self._h.close()

    def getFoveaWidth(self):
        return self.fovW

    def getFoveaHeight(self):
        return self.fovH

    def getFoveateLevel(self):
        return self.foveatelevel

    def setFoveaWidth(self, rows):
        self.fovW = rows

    def setFoveaHeight(self, cols):
        self.fovH = cols

    def setFoveated(self, fov):
        self.foveatedmatching = fov

    def match(self, rgbL, rgbR, fov=0):
        """-> [H, V, C], each (rows, cols) float32 (MatchGPULib.cpp:303-403)."""
        self.foveatedmatching = fov
        if fov:  # foveated matching + hierarchicalDisparity composite (MatchGPULib.cpp:355-360)
            return list(self._h.match_composite(rgbL, rgbR))
        return list(self._h.match_full(rgbL, rgbR))

    def initStack(self, rgbL, rgbR=None):
        rows, cols = np.asarray(rgbL).shape[:2]
        d = capi.query_dims(self._h.cfg, rows, cols, 1)
        self.setFoveaHeight(d.fov_height)
        self.setFoveaWidth(d.fov_width)
        return 0

    def matchStack(self, rgbL, rgbR):
        """-> disp[14][3]: per level [H, V, C]; levels below the fovea level are fovH x fovW (MatchGPULib.cpp:429-531)."""
        return self.matchStackPyramid(rgbL, rgbR)[0]

    def matchStackPyramid(self, rgbL, rgbR):
        """-> (disp[14][3], leftFov[14], rightFov[14]) (MatchGPULib.cpp:534-700)."""
        self.foveatedmatching = 1
        rows, cols = np.asarray(rgbL).shape[:2]
        _, _, _, self.fovW, self.fovH = self._h.match_foveated(rgbL, rgbR)
        levels = self._h.cfg.levels
        disp = [list(self._h.level_disparity(i, rows, cols, 1)) for i in range(levels)]
        fov = [self._h.fovea_images(i, rows, cols, 1) for i in range(levels)]
        return disp, [f[0] for f in fov], [f[1] for f in fov]
