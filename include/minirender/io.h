// minirender (B200 build) — image I/O on the render path (reference include/minirender/io.h:18-20,
// src/io.cpp:337-415). Mesh loaders (STL/OBJ/X3D) are outside the hot-path scope.
#ifndef MINIRENDER_B200_IO_H
#define MINIRENDER_B200_IO_H

#include "Scene.h"

namespace minirender {

// Binary P6 writer; each channel is (byte)clamp(v*255, 0, 255), i.e. truncation
// (reference src/io.cpp:358-361). "--" writes to stdout.
void savePPM(const asl::Array2<asl::Vec3>& image, const asl::String& filename);

// Binary P6 reader: '#' comments allowed in the header, maxval ignored, texel = rgb/255
// (reference src/io.cpp:367-415). Returns an empty image on failure.
asl::Array2<asl::Vec3> loadPPM(const asl::String& filename);

// The 8-bit quantiser alone (what parity on RGB is judged in).
void quantizeRGB8(const asl::Array2<asl::Vec3>& image, asl::byte* rgb8);

}
#endif
