// sbp_driver.cpp -- builds a Frame pair from plain arrays, runs the REFERENCE's own host code of
// ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) on it and hands the outcome back through a C entry point.
// TEST INFRASTRUCTURE (oracle/): produces the fixtures tests/golden/sbpref_*.npz (tools/make_golden_sbp.py).
// The function bodies are NOT in this repository: extract_slice.py cuts them out of the reference checkout by line range at
// build time (oracle/_ref/gen/*.inc); this file only supplies what surrounds them.
#include "sbp_stub.hpp"

float Frame::fx, Frame::fy, Frame::cx, Frame::cy, Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;

#include "frame_assign_grid.inc"          // void Frame::AssignFeaturesToGrid()                    src/Frame.cpp:464-479
#include "frame_features_in_area.inc"     // void Frame::GetFeaturesInArea(x, y, invzc, r, ...)    src/Frame.cpp:569-639
#include "frame_pos_in_grid.inc"          // bool Frame::PosInGrid(kp, posX, posY)                 src/Frame.cpp:696-706
#include "sbp_three_maxima.inc"           // void ORBmatcher::ComputeThreeMaxima(...)             src/ORBmatcher.cpp:2097-2138

// src/ORBmatcher.cpp:1315-1321 (signature and the two locals in front of the branch), then the live branch as cut out
int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)
{
    int nmatches = 0;
    bool use_gpu_ = true;
    if(!use_gpu_)
    {
    }
#include "sbp_live_branch.inc"            //  else { ... }                                        src/ORBmatcher.cpp:1647-1963
    return nmatches;
}

extern "C" {
// last frame: n_last keypoints; has_mp[i] != 0 <=> mvpMapPoints[i] != NULL (world position P[3][n_last], descriptor last_desc[i]),
// outlier[i] = mvbOutlier[i].  current frame: n_cur keypoints with occupied[j] != 0 <=> a map point with Observations() > 0.
// pose_cur / pose_last: 4x4 row-major Tcw.  Outputs: cur_match[j] = index i of the last-frame keypoint whose map point
// CurrentFrame.mvpMapPoints[j] ends up holding (-1: none or the pre-existing one), level_mode = 0 / 1 (bForward) / 2 (bBackward)
// as the function chose it.  Returns nmatches.
int jsref_search_by_projection(int n_last, const float* P, const unsigned char* has_mp, const unsigned char* outlier, const int* last_octave,
                               const float* last_angle, const unsigned char* last_desc, const float* pose_last, int n_cur,
                               const float* cur_x, const float* cur_y, const int* cur_octave, const float* cur_angle,
                               const float* cur_uright, const unsigned char* occupied, const unsigned char* cur_desc,
                               const float* pose_cur, float fx, float fy, float cx, float cy, float min_x, float max_x, float min_y,
                               float max_y, float mbf, float mb, float th, const float* scale_factors, int n_levels, int b_mono,
                               int check_orientation, int* cur_match, int* level_mode) {
    Frame::fx = fx; Frame::fy = fy; Frame::cx = cx; Frame::cy = cy;
    Frame::mnMinX = min_x; Frame::mnMaxX = max_x; Frame::mnMinY = min_y; Frame::mnMaxY = max_y;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (max_x - min_x);     // src/Frame.cpp:234-235
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (max_y - min_y);
    Frame last, cur;
    auto pose = [](const float* t) { cv::Mat m(4, 4); for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m.f(i, j) = t[4 * i + j]; return m; };
    last.mTcw = pose(pose_last); cur.mTcw = pose(pose_cur);
    last.N = n_last; cur.N = n_cur;
    last.mb = cur.mb = mb; last.mbf = cur.mbf = mbf;
    last.mvScaleFactors.assign(scale_factors, scale_factors + n_levels);
    cur.mvScaleFactors = last.mvScaleFactors;
    vector<MapPoint> pool(n_last + 1);
    last.mvpMapPoints.assign(n_last, nullptr);
    last.mvbOutlier.assign(n_last, false);
    last.mvKeys.resize(n_last); last.mvKeysUn.resize(n_last);
    for (int i = 0; i < n_last; ++i) {
        cv::KeyPoint k{}; k.octave = last_octave[i]; k.angle = last_angle[i]; k.class_id = -1;
        last.mvKeys[i] = last.mvKeysUn[i] = k;
        last.mvbOutlier[i] = outlier[i] != 0;
        if (has_mp[i]) {
            MapPoint& mp = pool[i];
            mp.x = P[i]; mp.y = P[n_last + i]; mp.z = P[2 * n_last + i];
            memcpy(mp.desc, last_desc + 32 * (size_t)i, 32);
            last.mvpMapPoints[i] = &mp;
        }
    }
    MapPoint& busy = pool[n_last];       // what an already-tracked keypoint of the current frame points at
    busy.nObs = 1;
    cur.mvpMapPoints.assign(n_cur, nullptr);
    cur.mvKeys.resize(n_cur); cur.mvKeysUn.resize(n_cur);
    cur.mvuRight.assign(cur_uright, cur_uright + n_cur);
    for (int j = 0; j < n_cur; ++j) {
        cv::KeyPoint k{}; k.pt.x = cur_x[j]; k.pt.y = cur_y[j]; k.octave = cur_octave[j]; k.angle = cur_angle[j]; k.class_id = -1;
        cur.mvKeys[j] = cur.mvKeysUn[j] = k;
        if (occupied[j]) cur.mvpMapPoints[j] = &busy;
    }
    cur.mDescriptors = cv::Mat::bytes(const_cast<unsigned char*>(cur_desc), n_cur, 32);
    cur.AssignFeaturesToGrid();
    ORBmatcher matcher;
    matcher.mbCheckOrientation = check_orientation != 0;
    const int n = matcher.SearchByProjection(cur, last, th, b_mono != 0);
    for (int j = 0; j < n_cur; ++j) {
        MapPoint* q = cur.mvpMapPoints[j];
        cur_match[j] = (q && q != &busy) ? (int)(q - pool.data()) : -1;
    }
    {   // the same expressions the function evaluates (src/ORBmatcher.cpp:1656-1669), reported for the fixture
        const cv::Mat Rcw = cur.mTcw.rowRange(0, 3).colRange(0, 3), tcw = cur.mTcw.rowRange(0, 3).col(3);
        const cv::Mat twc = -Rcw.t() * tcw;
        const cv::Mat tlc = last.mTcw.rowRange(0, 3).colRange(0, 3) * twc + last.mTcw.rowRange(0, 3).col(3);
        const bool fwd = tlc.at<float>(2) > cur.mb && !b_mono, bwd = -tlc.at<float>(2) > cur.mb && !b_mono;
        *level_mode = fwd ? 1 : bwd ? 2 : 0;
    }
    return n;
}
}
