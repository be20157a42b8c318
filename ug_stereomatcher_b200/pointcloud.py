"""Host-side mirror of the reconstruction half of the reference's point-cloud node
(class CdynamicCalibration, src/pointcloud/getPointCloud.cpp): same method names and argument
meaning, the per-pixel loops replaced by one K5 launch through the C-ABI (ugsm_cloud_*).

Only the triangulated clouds are provided (doReconstructionRGB, PC:741-782; doReconstructionRGB_FOV,
PC:697-740).  The "resized" clouds (PC:784-884) resample the range map with OpenCV's bicubic filter and
are out of scope.  ROS plumbing (subscribers, synchronisers, publishers) stays in the node.
"""
import numpy as np

from . import capi


class CdynamicCalibration:
    def __init__(self, handle=None, sampling=1, foveated=0, srcLevel=0):
        self._own = handle is None
        self.h = handle if handle is not None else capi.Handle()
        self.sampling = sampling  # PC:1196: 1; "for the non fovea version change this to 5"
        self.foveated = foveated
        self.srcLevel = srcLevel  # PC:551-566: stack level, 0 = finest fovea
        self.destLevel = 0        # PC:569
        self.cam = None
        self.imgL = None

    def close(self):
        if self._own:
            self.h.close()

    def getCameraInfo(self, P_left, P_right):
        """PC:1119-1178: P of both CameraInfo messages, 12 doubles each, row-major 3x4."""
        try:
            self.cam = capi.CameraPair.of(P_left, P_right)
        except Exception:
            return False
        return True

    def getImages(self, imgL_bgr):
        """PC:236-282 keeps the left image as BGR8."""
        self.imgL = np.ascontiguousarray(imgL_bgr, np.uint8)

    def _ready(self):
        if self.cam is None or self.imgL is None:
            raise capi.UgsmError(capi.ERR_INVALID, "camera info and left image must be set first")

    def doReconstructionRGB(self, dispX, dispY, stride=4):
        """-> (n, stride) float32: x, y, z, rgb-bits per point, column-major point order (PC:746-779)."""
        self._ready()
        return self.h.cloud_full(self.cam, dispX, dispY, self.imgL, bgr=True, sampling=self.sampling, stride=stride)

    def doReconstructionRGB_FOV(self, fdispX, fdispY, stride=4):
        """fdispX / fdispY: the fovea stacks ((7 * fovH) x fovW).  PC:697-740."""
        self._ready()
        fovea_level = self.h.cfg.fovea_level  # foveaLevel, PC:56
        fovH = fdispX.shape[0] // fovea_level  # PC:706
        return self.h.cloud_foveated(self.cam, fdispX, fdispY, fovH, self.srcLevel, self.imgL, bgr=True,
                                     sampling=self.sampling, stride=stride)

    def fromLastMatch(self, rows, cols, src_level=-1, stride=4):
        """Not in the reference: the cloud of the pair last matched on the same handle, straight from the
        planes in device memory (the matcher and this class then share one ugsm handle)."""
        if self.cam is None:
            raise capi.UgsmError(capi.ERR_INVALID, "camera info must be set first")
        return self.h.cloud_from_match(self.cam, rows, cols, src_level=src_level, sampling=self.sampling,
                                       stride=stride)
