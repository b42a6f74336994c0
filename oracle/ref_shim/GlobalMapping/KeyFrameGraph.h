// oracle/ref_shim/GlobalMapping/KeyFrameGraph.h -- TEST INFRASTRUCTURE.
// The reference's DepthMap.cpp:33, DepthMapPixelHypothesis.cpp:22 and TrackingReference.cpp:24 include
// GlobalMapping/KeyFrameGraph.h (which pulls in g2o) without using anything from it; the pose graph is out of scope
// (SURVEY.md section 2), so the oracle/_ref build sees this empty declaration instead.
#pragma once
namespace lsd_slam { class KeyFrameGraph; }
