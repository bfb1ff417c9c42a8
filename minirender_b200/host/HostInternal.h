// Shared between the host sources of the library; not part of the API.
#ifndef MINIRENDER_B200_HOST_INTERNAL_H
#define MINIRENDER_B200_HOST_INTERNAL_H

#include <atomic>

namespace minirender {

// Counts the in-place geometry edits the library performs itself (TriMesh::applyTransform): a Renderer whose mirror of
// the scene in HBM was uploaded under another count uploads again, whatever its sampled fingerprints say.
std::atomic<unsigned>& geometryEpoch();

}
#endif
