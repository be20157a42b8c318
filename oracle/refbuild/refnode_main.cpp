// refnode_main.cpp -- TEST INFRASTRUCTURE.  The reference's matcher node (src/gpu_matcher/UG_GPU_matcher.cpp),
// compiled UNMODIFIED, running on this repo's library: the only thing that changes is which MatchGPULib.h
// its `#include "MatchGPULib.h"` resolves to (include/MatchGPULib.h of this repo; the reference's own header
// next to the node source is switched off through its include guard, which is what replacing that file does).
// ROS / OpenCV / cv_bridge are the inert stand-ins of stubs_pc/; messages the node publishes are captured
// and written out so that tests/test_gpu_node.py can compare them with the C-ABI results.
//
//   refnode <in.bin> <out.bin> [node args...]      (node args: argv[2] = fovea level, as the node takes it)
//   in : int32 mode, rows, cols; uint8 rgbL[rows*cols*3], rgbR[rows*cols*3]
//        mode bit 0 = the node's `foveated` parameter; bit 1 = go through the service callback disparitySrv
//        (UG_GPU_matcher.cpp:497-690, srv/GetDisparitiesGPU.srv) instead of the topic callback mainRoutine
//   out: for each topic, then for each response field of the service, in a fixed order:
//        int32 rows, cols; float data[rows*cols]   (rows = 0: not published / not filled)
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "pc_stubs.h"

#define UGSM_WITH_CV_BRIDGE 1
#include <MatchGPULib.h>  // this repo's facade
#define _MATCHGPULIB_H    // the reference header next to the node source is now an empty include

#define main ugsm_unused_reference_main
#include "src/gpu_matcher/UG_GPU_matcher.cpp"
#undef main

static void rd(FILE* f, void* p, size_t n) {
    if (fread(p, 1, n, f) != n) {
        fprintf(stderr, "refnode: short read\n");
        exit(2);
    }
}

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: refnode in.bin out.bin [node args]\n");
        return 2;
    }
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t hd[3];
    rd(f, hd, sizeof hd);
    const int foveated = hd[0] & 1, service = (hd[0] >> 1) & 1, rows = hd[1], cols = hd[2];
    sensor_msgs::Image* imL = new sensor_msgs::Image;
    sensor_msgs::Image* imR = new sensor_msgs::Image;
    imL->payload = cv::Mat(rows, cols, CV_8UC3);
    imR->payload = cv::Mat(rows, cols, CV_8UC3);
    rd(f, imL->payload.data, (size_t)rows * cols * 3);
    rd(f, imR->payload.data, (size_t)rows * cols * 3);
    fclose(f);
    imL->height = imR->height = rows;
    imL->width = imR->width = cols;
    imL->encoding = imR->encoding = "rgb8";

    ugsm_stub::params()["foveated"] = foveated;
    // the node hands its own argc/argv to MatchGPULib (UG_GPU_matcher.cpp:160): argv[0] = program, then the rest
    std::vector<char*> nargv;
    nargv.push_back(argv[0]);
    for (int i = 3; i < argc; i++) nargv.push_back(argv[i]);
    GPU_matcher node((int)nargv.size(), nargv.data());
    ug_stereomatcher::GetDisparitiesGPU::Request req;
    ug_stereomatcher::GetDisparitiesGPU::Response rsp;
    if (service) {
        req.imL = *imL;
        req.imR = *imR;
        if (!node.disparitySrv(req, rsp)) {
            fprintf(stderr, "refnode: disparitySrv returned false\n");
            return 3;
        }
    } else {
        node.mainRoutine(sensor_msgs::ImageConstPtr(imL), sensor_msgs::ImageConstPtr(imR));
    }

    static const char* topics[] = {"output_disparityH", "output_disparityV", "output_disparityC", "output_stackH",
                                   "output_stackV",     "output_stackC",     "output_stackL_pyramid",
                                   "output_stackR_pyramid"};
    FILE* o = fopen(argv[2], "wb");
    if (!o) return 2;
    for (const char* t : topics) {
        int32_t dims[2] = {0, 0};
        auto it = ugsm_stub::published().find(t);
        if (it != ugsm_stub::published().end()) {
            dims[0] = it->second.rows;
            dims[1] = it->second.cols;
        }
        fwrite(dims, sizeof dims, 1, o);
        if (dims[0] > 0) fwrite(it->second.data, 4, (size_t)dims[0] * dims[1], o);
    }
    // GetDisparitiesGPU response: dispH, dispV, dispC (.image), fdispH, fdispV, fdispC (.image_stack)
    const cv::Mat* fields[6] = {&rsp.dispH.image.payload,       &rsp.dispV.image.payload,       &rsp.dispC.image.payload,
                                &rsp.fdispH.image_stack.payload, &rsp.fdispV.image_stack.payload, &rsp.fdispC.image_stack.payload};
    for (const cv::Mat* mt : fields) {
        int32_t dims[2] = {mt->data ? mt->rows : 0, mt->data ? mt->cols : 0};
        fwrite(dims, sizeof dims, 1, o);
        if (dims[0] > 0) fwrite(mt->data, 4, (size_t)dims[0] * dims[1], o);
    }
    fclose(o);
    return 0;
}
