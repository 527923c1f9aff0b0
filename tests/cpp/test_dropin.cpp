// test_dropin.cpp -- the reference's call sites, compiled UNCHANGED in form against include/filter.h / include/util.h
// (the drop-in headers) and checked bit-for-bit against the CPU oracle (oracle/libdmc_oracle.so, loaded with dlopen
// so that no oracle symbol is linked into anything the product ships).
//
// cv::Mat is provided by oracle/refshim/minicv.hpp here (no OpenCV C++ headers exist in this image); with a real
// OpenCV the same file builds with -I<opencv include> instead of -Ioracle/refshim.
//
// Call sites mirrored: simpleTest() main.cpp:507-539, pointcloudTest() main.cpp:255-321,
// binalyWeightedRangeFilterTest() main.cpp:470-505, and the commented boundaryReconstructionFilter call main.cpp:308.
#include "filter.h"
#include "filter_ext.h"
#include "util.h"
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <vector>

#define FOCUS 75.0
#define BASELINE 575.0
#define AMP_DISP 2.6

static void* g_orc = 0;
template <class F> static F sym(const char* name) { F f = (F)dlsym(g_orc, name); if (!f) { fprintf(stderr, "missing %s\n", name); exit(2); } return f; }
static int g_fail = 0;

static bool same_bits(const Mat& a, const void* b, size_t bytes, bool is_float) {
    if (!is_float) return memcmp(a.data, b, bytes) == 0;
    const float* x = (const float*)a.data; const float* y = (const float*)b;
    for (size_t i = 0; i < bytes / 4; i++) { if (x[i] != x[i] && y[i] != y[i]) continue; if (memcmp(x + i, y + i, 4)) return false; }
    return true;
}
#define EXPECT(cond, what) do { if (!(cond)) { printf("FAIL %s\n", what); g_fail++; } else printf("ok   %s\n", what); } while (0)

static Mat make_disp(int rows, int cols, unsigned seed) {        // piecewise-constant blocks + noise, never 0
    Mat m(rows, cols, CV_8U); unsigned s = seed * 2654435761u + 12345u;
    std::vector<int> blk(((rows + 7) / 8) * ((cols + 7) / 8));
    for (size_t i = 0; i < blk.size(); i++) { s = s * 1664525u + 1013904223u; blk[i] = 20 + (s >> 24) % 200; }
    for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) {
        s = s * 1664525u + 1013904223u; int v = blk[(y / 8) * ((cols + 7) / 8) + x / 8] + (int)((s >> 28) % 13) - 6;
        m.at<uchar>(y, x) = (uchar)std::min(255, std::max(1, v)); }
    return m;
}

int main(int argc, char** argv) {
    const char* orc = argc > 1 ? argv[1] : "oracle/libdmc_oracle.so";
    g_orc = dlopen(orc, RTLD_NOW | RTLD_LOCAL);
    if (!g_orc) { fprintf(stderr, "cannot load %s: %s\n", orc, dlerror()); return 2; }
    typedef int (*pfs_t)(const uchar*, uchar*, int, int, int, int, int, int, int, int);
    typedef int (*d32_t)(const uchar*, float*, int, int, double, double, double, int, int, int, int, float, int);
    typedef int (*d16_t)(const uchar*, ushort*, int, int, double, double, double, int, int, int, int, float, int);
    typedef int (*dsp_t)(const uchar*, ushort*, int, int, int, int, int, int, float, int);
    typedef int (*bw_t)(const void*, void*, int, int, int, int, int, float, int);
    typedef int (*brf_t)(const void*, void*, int, int, int, int, int, float, float, float);
    typedef int (*mm_t)(const void*, void*, int, int, int, int);
    typedef int (*cv_t)(const void*, void*, int, int, float, float, float);
    typedef int (*xyz_t)(const void*, float*, int, int, int, double);
    typedef int (*jbw_t)(const uchar*, const uchar*, uchar*, int, int, int, int, int, float);
    pfs_t o_pfs = sym<pfs_t>("orc_post_filter_set"); d32_t o_d32 = sym<d32_t>("orc_filter_disp8u_depth32f");
    d16_t o_d16 = sym<d16_t>("orc_filter_disp8u_depth16u"); dsp_t o_dsp = sym<dsp_t>("orc_filter_disp8u_disp32f");
    bw_t o_bw = sym<bw_t>("orc_bwrf"); brf_t o_brf = sym<brf_t>("orc_brf"); mm_t o_mm = sym<mm_t>("orc_blur_remove_minmax");
    cv_t o_d2d = sym<cv_t>("orc_depth32f2disp8u"); xyz_t o_xyz = sym<xyz_t>("orc_reproject_xyz");
    jbw_t o_jbw = sym<jbw_t>("orc_joint_bwrf");
    typedef int (*ren_t)(const uchar*, const float*, int, int, const double*, const double*, const double*, int, int, uchar*, float*);
    typedef int (*hole_t)(const uchar*, uchar*, int, int);
    ren_t o_ren = sym<ren_t>("orc_project_image_serial"); hole_t o_hole = sym<hole_t>("orc_fill_small_hole");

    for (int trial = 0; trial < 3; trial++) {
        const int rows = trial == 0 ? 480 : trial == 1 ? 131 : 64, cols = trial == 0 ? 640 : trial == 1 ? 150 : 641;
        Mat disp8coded = make_disp(rows, cols, 7 + trial);
        const size_t n = (size_t)rows * cols;

        // --- simpleTest(), main.cpp:523-526 -------------------------------------------------------------------------
        Mat disp8filtered;//post filtered disparity map
        PostFilterSet pfs;//class of our post filter set
        pfs(disp8coded,disp8filtered,2,1,3,5,10);//post filter set
        std::vector<uchar> w8(n); o_pfs(disp8coded.data, w8.data(), rows, cols, 2, 1, 3, 5, 10, 0);
        EXPECT(disp8filtered.type() == CV_8U && same_bits(disp8filtered, w8.data(), n, false), "simpleTest: pfs(disp8coded,disp8filtered,2,1,3,5,10)");

        // --- pointcloudTest(), main.cpp:303-304, :321 ---------------------------------------------------------------
        Mat disp = disp8coded, depthF, dshow, xyz;
        double focus = FOCUS, baseline = BASELINE, amp = AMP_DISP; int mr = 1, gr = 0, br = 1, dr = 3, thresh = 65;
        pfs.filterDisp8U2Depth32F(disp,depthF,focus,baseline,amp,mr,gr,br,dr,(float)thresh,FULL_KERNEL);
        std::vector<float> w32(n); o_d32(disp.data, w32.data(), rows, cols, focus, baseline, amp, mr, gr, br, dr, (float)thresh, 0);
        EXPECT(depthF.type() == CV_32F && same_bits(depthF, w32.data(), n * 4, true), "pointcloudTest: pfs.filterDisp8U2Depth32F(...)");
        depth32F2disp8U(depthF,dshow,(float)(focus*baseline),(float)amp,0.f);
        std::vector<uchar> wd(n); o_d2d(w32.data(), wd.data(), rows, cols, (float)(focus * baseline), (float)amp, 0.f);
        EXPECT(same_bits(dshow, wd.data(), n, false), "pointcloudTest: depth32F2disp8U(depthF,dshow,...)");
        reprojectXYZ(depthF,xyz,510.0);
        std::vector<float> wx(n * 3); o_xyz(w32.data(), wx.data(), rows, cols, CV_32F, 510.0);
        EXPECT(xyz.type() == CV_32FC3 && same_bits(xyz, wx.data(), n * 12, true), "pointcloudTest: reprojectXYZ(depthF,xyz,focal_length)");

        // --- pointcloudTest(), main.cpp:132-136, :322-355: render from another viewpoint (isSub), fill the small holes -----
        {
            float focal_length=510.f;
            Mat k = Mat::eye(3,3,CV_64F)*focal_length;
            k.at<double>(0,2)=(cols-1)*0.5;
            k.at<double>(1,2)=(rows-1)*0.5;
            k.at<double>(2,2)=1.0;
            Mat R = Mat::eye(3,3,CV_64F);
            Mat t = Mat::zeros(3,1,CV_64F);
            t.at<double>(0,0)=180.0; t.at<double>(1,0)=-90.0; t.at<double>(2,0)=350.0;
            Mat dispC(rows, cols, CV_8UC3), destImage, nodist, nomask;
            for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) { uchar v = dshow.at<uchar>(y, x); uchar* c = dispC.ptr<uchar>(y) + 3 * x; c[0] = c[1] = c[2] = v; }   // cvtColor(dshow,dispC,CV_GRAY2BGR) :346
            projectImagefromXYZ(dispC,destImage,xyz,R,t,k,nodist,nomask,true);
            std::vector<uchar> wim(n * 3); std::vector<float> wz(n);
            const double Rd[9] = {1,0,0, 0,1,0, 0,0,1}, td[3] = {180.0, -90.0, 350.0}, Kd[9] = {510.0,0,(cols-1)*0.5, 0,510.0,(rows-1)*0.5, 0,0,1.0};
            o_ren(dispC.data, wx.data(), rows, cols, Rd, td, Kd, 1, 1, wim.data(), wz.data());
            EXPECT(destImage.type() == CV_8UC3 && same_bits(destImage, wim.data(), n * 3, false), "pointcloudTest: projectImagefromXYZ(dispC,destImage,xyz,R,t,k,Mat(),Mat(),true)  [Intel host: rcp table]");
            vector<Point2f> pt(n); Mat zbuf = Mat::zeros(rows, cols, CV_32F), dest2;
            projectImagefromXYZ(dispC,dest2,xyz,R,t,k,nodist,nomask,true,pt,zbuf);
            EXPECT(same_bits(dest2, wim.data(), n * 3, false) && same_bits(zbuf, wz.data(), n * 4, true), "projectImagefromXYZ(..., pt, depth) overload");
            fillSmallHole(destImage,destImage);
            o_hole(wim.data(), wim.data(), rows, cols);
            EXPECT(same_bits(destImage, wim.data(), n * 3, false), "pointcloudTest: fillSmallHole(destImage,destImage)");
            Mat planes; splitBGRLineInterleave(dispC, planes);
            bool okp = planes.rows == 3 * rows && planes.cols == cols && planes.type() == CV_8U;
            for (int y = 0; okp && y < rows; y++) for (int x = 0; x < cols; x++) for (int c = 0; c < 3; c++) if (planes.at<uchar>(3 * y + c, x) != dispC.ptr<uchar>(y)[3 * x + c]) { okp = false; break; }
            EXPECT(okp, "splitBGRLineInterleave(src, dest)");
        }

        // --- the other two entry points -----------------------------------------------------------------------------
        Mat d16, dd16; std::vector<ushort> w16(n);
        pfs.filterDisp8U2Depth16U(disp, d16, focus, baseline, amp, mr, gr, br, dr, (float)thresh);
        o_d16(disp.data, w16.data(), rows, cols, focus, baseline, amp, mr, gr, br, dr, (float)thresh, 0);
        EXPECT(d16.type() == CV_16U && same_bits(d16, w16.data(), n * 2, false), "pfs.filterDisp8U2Depth16U(...)");
        pfs.filterDisp8U2Disp32F(disp, dd16, mr, gr, br, dr, 10.f);
        o_dsp(disp.data, w16.data(), rows, cols, mr, gr, br, dr, 10.f, 0);
        EXPECT(dd16.type() == CV_16U && same_bits(dd16, w16.data(), n * 2, false), "pfs.filterDisp8U2Disp32F(...)");

        // --- binalyWeightedRangeFilterTest(), main.cpp:477-495 ------------------------------------------------------
        Mat dest1, dest2, input; Size ksize = Size(5,5); float thr = 8;
        disp8coded.convertTo(input,CV_32F);
        binalyWeightedRangeFilter(input, dest1, ksize, thr, FULL_KERNEL);
        binalyWeightedRangeFilter(input, dest2, ksize, thr, FULL_KERNEL_PAIR);
        std::vector<float> wb(n); o_bw(input.data, wb.data(), rows, cols, CV_32F, 5, 5, thr, 0);
        EXPECT(same_bits(dest1, wb.data(), n * 4, true), "binalyWeightedRangeFilter(input, dest1, ksize, thresh, FULL_KERNEL)");
        EXPECT(same_bits(dest2, wb.data(), n * 4, true), "binalyWeightedRangeFilter(..., FULL_KERNEL_PAIR) == FULL_KERNEL result (documented)");
        Mat keep = disp8coded.clone(), untouched = keep.clone();
        binalyWeightedRangeFilter(disp8coded, keep, ksize, 10.f, FULL_KERNEL_PAIR);       // 8U + PAIR: silent no-op
        EXPECT(same_bits(keep, untouched.data, n, false), "8U + FULL_KERNEL_PAIR leaves dst untouched");

        // --- main.cpp:308 (commented competitor): boundaryReconstructionFilter(disp,disp,Size(13,13),1,1,1), in place ----
        Mat b = disp8coded.clone(); std::vector<uchar> wbr(n);
        o_brf(disp8coded.data, wbr.data(), rows, cols, CV_8U, 13, 13, 1.f, 1.f, 1.f);
        boundaryReconstructionFilter(b,b,Size(13,13),1,1,1);
        EXPECT(same_bits(b, wbr.data(), n, false), "boundaryReconstructionFilter(disp,disp,Size(13,13),1,1,1)");

        // --- blurRemoveMinMax, in place like the chain uses it (postFilterSet.cpp:25) --------------------------------
        Mat m = disp8coded.clone(); std::vector<uchar> wm(n);
        o_mm(disp8coded.data, wm.data(), rows, cols, CV_8U, 3);
        blurRemoveMinMax(m,m,3);
        EXPECT(same_bits(m, wm.data(), n, false), "blurRemoveMinMax(buff,buff,minmax_r)");

        // --- extension (include/filter_ext.h, no reference call site): range filter of the depth map guided by a colour image ----
        Mat guide(rows, cols, CV_8UC3), joint; std::vector<uchar> wj(n);
        for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) { uchar v = disp8coded.at<uchar>(y, x); uchar* g = guide.ptr<uchar>(y) + 3 * x;
            g[0] = (uchar)(255 - v); g[1] = (uchar)((v * 7) & 0xF0); g[2] = (uchar)(v / 2 + ((x ^ y) & 3)); }
        jointBinalyWeightedRangeFilter(disp8coded, guide, joint, Size(11,11), 30.f);
        o_jbw(disp8coded.data, guide.data, wj.data(), rows, cols, 3, 11, 11, 30.f);
        EXPECT(joint.type() == CV_8U && same_bits(joint, wj.data(), n, false), "jointBinalyWeightedRangeFilter(disp, colour guide, dst, Size(11,11), 30)");

        // --- extension: blurRemoveMinMax + boundaryReconstructionFilter in one kernel == the two reference calls ----------
        Mat fused; std::vector<uchar> wmm(n), wfb(n);
        o_mm(disp8coded.data, wmm.data(), rows, cols, CV_8U, 3);
        o_brf(wmm.data(), wfb.data(), rows, cols, CV_8U, 13, 13, 1.f, 1.f, 1.f);
        minmaxBoundaryReconstructionFilter(disp8coded, fused, 3, Size(13,13), 1, 1, 1);
        EXPECT(fused.type() == CV_8U && same_bits(fused, wfb.data(), n, false), "minmaxBoundaryReconstructionFilter(disp, dst, 3, Size(13,13), 1,1,1)");

        // --- error convention: CV_Assert(src.type()==dst.type()) -> cv::Exception -------------------------------------
        bool threw = false; Mat wrong(rows, cols, CV_32F);
        try { binalyWeightedRangeFilter(disp8coded, wrong, ksize, 10.f, FULL_KERNEL); } catch (const std::exception&) { threw = true; }
        EXPECT(threw, "type mismatch throws");
    }
    printf("%s (%d failures)\n", g_fail ? "FAILED" : "ALL OK", g_fail);
    return g_fail ? 1 : 0;
}
