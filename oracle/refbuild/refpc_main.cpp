// refpc_main.cpp -- TEST INFRASTRUCTURE.  Runs the reference's point-cloud reconstruction
// (src/pointcloud/getPointCloud.cpp, compiled UNMODIFIED through stubs_pc/) on a binary input file
// and writes the points it produces, so the oracle's restatement (oracle/ugsm_oracle.c,
// ugsmo_pointcloud_*) can be pinned against the real code.  CPU only.
//
//   refpc <in.bin> <out.bin>
//   in : int32 mode(0 full, 1 foveated; 2 / 3: the range maps of the same two inputs, i.e. getRangePoint /
//        getRangePointFOV (getPointCloud.cpp:951-1019) at every pixel -- out: int64 n; n x float, row-major),
//        rows, cols (of the disparity planes / of the fovea stack),
//        im_rows, im_cols, sampling, src_level; double P1[12], P2[12]; float dX[rows*cols],
//        dY[rows*cols]; uint8 bgr[im_rows*im_cols*3]
//   out: int64 n; n x {float x, y, z; uint32 rgb}
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "pc_stubs.h"

// the node class keeps its reconstruction methods private and the file has its own main()
#define private public
#define main ugsm_unused_reference_main
#include "src/pointcloud/getPointCloud.cpp"
#undef main
#undef private

static void rd(FILE* f, void* p, size_t n) {
    if (fread(p, 1, n, f) != n) {
        fprintf(stderr, "refpc: short read\n");
        exit(2);
    }
}

int main(int argc, char** argv) {
    if (argc != 3) {
        fprintf(stderr, "usage: refpc in.bin out.bin\n");
        return 2;
    }
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t hd[7];
    rd(f, hd, sizeof hd);
    const int mode = hd[0], rows = hd[1], cols = hd[2], im_rows = hd[3], im_cols = hd[4], sampling = hd[5],
              src_level = hd[6];
    double P1[12], P2[12];
    rd(f, P1, sizeof P1);
    rd(f, P2, sizeof P2);
    cv::Mat dX(rows, cols, CV_32F), dY(rows, cols, CV_32F), img(im_rows, im_cols, CV_8UC3);
    rd(f, dX.data, (size_t)rows * cols * 4);
    rd(f, dY.data, (size_t)rows * cols * 4);
    rd(f, img.data, (size_t)im_rows * im_cols * 3);
    fclose(f);

    // stdout carries the node's progress bar; silence it
    std::cout.setstate(std::ios_base::failbit);

    CdynamicCalibration node;
    node.sampling = sampling;
    node.save_cloud = false;
    node.imgL = img;
    node.set_hInit(im_rows);
    node.set_wInit(im_cols);
    cv::Mat p1(3, 4, CV_64F), p2(3, 4, CV_64F);
    for (int i = 0; i < 12; i++) {
        p1.at<double>(i / 4, i % 4) = P1[i];
        p2.at<double>(i / 4, i % 4) = P2[i];
    }
    node.P1_ = p1;
    node.P2_ = p2;
    if (mode == 2 || mode == 3) {  // range maps: the loops of doReconstruction_resized / doReconstructionFOV_resized
        std::vector<float> range;      // (getPointCloud.cpp:730-751, 806-834) up to, not including, cv::resize
        int rr, rc;
        if (mode == 2) {
            node.foveated = 0;
            node.dispX = dX;
            node.dispY = dY;
            rr = rows, rc = cols;
            range.resize((size_t)rr * rc);
            for (int ii = 0; ii < rc; ii++)
                for (int jj = 0; jj < rr; jj++) range[(size_t)jj * rc + ii] = node.getRangePoint(ii, jj);
        } else {
            node.foveated = 1;
            node.srcLevel = src_level;
            node.destLevel = 0;
            node.fdispX = dX;
            node.fdispY = dY;
            node.leftMargin_offset = node.left_marginOf_in(node.srcLevel, node.destLevel);
            node.upperMargin_offset = node.upper_marginOf_in(node.srcLevel, node.destLevel);
            node.fovH = node.fdispX.rows / foveaLevel;  // a macro of the reference source (getPointCloud.cpp:63)
            node.fovW = node.fdispX.cols;
            rr = node.fovH, rc = node.fovW;
            range.resize((size_t)rr * rc);
            for (int ii = 0; ii < rc; ii++)
                for (int jj = 0; jj < rr; jj++) range[(size_t)jj * rc + ii] = node.getRangePointFOV(ii, jj);
        }
        FILE* o = fopen(argv[2], "wb");
        if (!o) return 2;
        int64_t n = (int64_t)range.size();
        fwrite(&n, sizeof n, 1, o);
        fwrite(range.data(), 4, range.size(), o);
        fclose(o);
        return 0;
    }
    pcl::PointCloud<pcl::PointXYZRGB>::Ptr cloud;
    if (mode == 0) {
        node.foveated = 0;
        node.dispX = dX;
        node.dispY = dY;
        cloud = node.doReconstructionRGB();
    } else {
        node.foveated = 1;
        node.srcLevel = src_level;
        node.destLevel = 0;
        node.fdispX = dX;
        node.fdispY = dY;
        cloud = node.doReconstructionRGB_FOV();
    }
    FILE* o = fopen(argv[2], "wb");
    if (!o) return 2;
    int64_t n = (int64_t)cloud->points.size();
    fwrite(&n, sizeof n, 1, o);
    for (int64_t i = 0; i < n; i++) {
        const pcl::PointXYZRGB& p = cloud->points[i];
        float rec[4] = {p.x, p.y, p.z, p.rgb};
        fwrite(rec, sizeof rec, 1, o);
    }
    fclose(o);
    return 0;
}
