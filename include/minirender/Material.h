// minirender (B200 build) — surface description (API of reference include/minirender/Scene.h:38-45,
// defaults of src/Scene.cpp:65-71: diffuse (0.7,0.7,0.9), specular 0.8 grey, no emission,
// shininess 12, opacity 1).
#ifndef MINIRENDER_B200_MATERIAL_H
#define MINIRENDER_B200_MATERIAL_H

#include <asl/Array2.h>
#include <asl/String.h>
#include <asl/Vec3.h>

namespace minirender {

struct Material
{
	asl::Vec3 diffuse;   // multiplied by (n.l / |n| + ambient); replaced by the texel when textured
	asl::Vec3 specular;  // multiplied by the Blinn-Phong term; skipped entirely when shininess == 0
	asl::Vec3 emissive;  // added unconditionally (the only term when lighting is off)
	float shininess;
	float opacity;       // carried for API compatibility; like the reference, the renderer ignores it
	asl::Array2<asl::Vec3> texture; // float RGB, rows x cols, nearest-texel, wrapped by fract(); empty = untextured
	asl::String textureName;

	Material();
};

}
#endif
