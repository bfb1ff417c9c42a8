// Mesh file formats on either side of the render path: STL (ASCII + binary),
// OBJ + MTL, X3D, and the two exporters saveSTL / saveXYZ. Written against plain stdio / std::string
// (the reference uses asl::TextFile / Xml / Dic / Path, reference src/io.cpp:16-335, src/x3d.cpp);
// each function follows the reference's parsing rules, cited below, so that the TriMesh arrays a
// file produces — and therefore the frames rendered from them — are the same.
//
// Deliberate, documented differences (all on inputs the reference handles by accident):
//  * one mesh per OBJ material is emitted in order of first use; the reference iterates an asl::Dic
//    (hash order, unspecified). Only the submission order of exactly coincident surfaces depends on it.
//  * an X3D <Inline> is read by its own reader; the reference re-uses one reader object and thereby
//    overwrites the document it is still iterating (x3d.cpp:169,176-180).
#include <minirender/io.h>

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace asl;

namespace minirender {

namespace {

typedef std::vector<std::string> Strings;

bool readFile(const std::string& name, std::string& out)
{
	FILE* f = fopen(name.c_str(), "rb");
	if (!f)
		return false;
	char buf[1 << 16];
	size_t n;
	out.clear();
	while ((n = fread(buf, 1, sizeof(buf), f)) > 0)
		out.append(buf, n);
	fclose(f);
	return true;
}

// asl::String::split(): whitespace-separated, empty tokens dropped
Strings splitWs(const std::string& s)
{
	Strings out;
	size_t i = 0;
	while (i < s.size())
	{
		while (i < s.size() && isspace((unsigned char)s[i]))
			i++;
		size_t j = i;
		while (j < s.size() && !isspace((unsigned char)s[j]))
			j++;
		if (j > i)
			out.push_back(s.substr(i, j - i));
		i = j;
	}
	return out;
}

// asl::String::split(sep): empty tokens kept ("1//3" -> "1", "", "3")
Strings splitChar(const std::string& s, char sep)
{
	Strings out;
	size_t i = 0;
	for (;;)
	{
		size_t j = s.find(sep, i);
		if (j == std::string::npos)
		{
			out.push_back(s.substr(i));
			break;
		}
		out.push_back(s.substr(i, j - i));
		i = j + 1;
	}
	return out;
}

float toFloat(const std::string& s) { return strtof(s.c_str(), 0); }
int toInt(const std::string& s) { return (int)strtol(s.c_str(), 0, 10); }

std::vector<float> floatsOf(const std::string& s)
{
	std::vector<float> v;
	const char* p = s.c_str();
	for (;;)
	{
		while (*p && (isspace((unsigned char)*p) || *p == ','))
			p++;
		if (!*p)
			break;
		char* e;
		const float f = strtof(p, &e);
		if (e == p)
			break;
		v.push_back(f);
		p = e;
	}
	return v;
}

std::vector<int> intsOf(const std::string& s)
{
	std::vector<int> v;
	const char* p = s.c_str();
	for (;;)
	{
		while (*p && (isspace((unsigned char)*p) || *p == ','))
			p++;
		if (!*p)
			break;
		char* e;
		const long k = strtol(p, &e, 10);
		if (e == p)
			break;
		v.push_back((int)k);
		p = e;
	}
	return v;
}

std::string directoryOf(const std::string& path)
{
	const size_t k = path.find_last_of("/\\");
	return k == std::string::npos ? std::string(".") : path.substr(0, k);
}

std::string noExt(const std::string& path)
{
	const size_t slash = path.find_last_of("/\\");
	const size_t dot = path.find_last_of('.');
	if (dot == std::string::npos || (slash != std::string::npos && dot < slash))
		return path;
	return path.substr(0, dot);
}

bool hasExtension(const std::string& path, const char* ext)
{
	const size_t dot = path.find_last_of('.');
	if (dot == std::string::npos)
		return false;
	std::string e = path.substr(dot + 1);
	for (size_t i = 0; i < e.size(); i++)
		e[i] = (char)tolower((unsigned char)e[i]);
	return e == ext;
}

Array<int> toArray(const std::vector<int>& v)
{
	Array<int> a;
	a.reserve((int)v.size());
	for (size_t i = 0; i < v.size(); i++)
		a << v[i];
	return a;
}

// Flat normals for a mesh that came without any (io.cpp:316-329, x3d.cpp:137-149):
// n = ((b - a) ^ (c - a)).normalized(), one per triangle.
void flatNormals(TriMesh* mesh)
{
	mesh->normalsI.clear();
	for (int i = 0, j = 0; i + 2 < mesh->indices.length(); i += 3, j++)
	{
		const Vec3 a = mesh->vertices[mesh->indices[i]];
		const Vec3 b = mesh->vertices[mesh->indices[i + 1]];
		const Vec3 c = mesh->vertices[mesh->indices[i + 2]];
		const Vec3 n = ((b - a) ^ (c - a)).normalized();
		mesh->normals << n;
		mesh->normalsI << j << j << j;
	}
}

}

// Fan triangulation of polygons separated by -1 (reference src/x3d.cpp:17-33): for a polygon
// starting at i, emits (i, j, j+1) until either j or j+1 is a terminator, then skips past it.
Array<int> triangulateIndices(const Array<int>& indices)
{
	Array<int> tris;
	tris.reserve(indices.length());
	int i = 0, j = 0;
	const int n = indices.length();
	while (i < n)
	{
		for (j = i + 1; j < n - 1; j++)
		{
			if (indices[j] == -1 || indices[j + 1] == -1)
				break;
			tris << indices[i] << indices[j] << indices[j + 1];
		}
		i = j + 2;
	}
	return tris;
}

// ------------------------------------------------------------------------------------------
// STL (reference src/io.cpp:16-111, 134-156)
// ------------------------------------------------------------------------------------------
static Shared<TriMesh> loadSTLa(const std::string& text)
{
	// token stream: "solid" first; "normal"/"vertex" consume the rest of their line (io.cpp:27-58)
	size_t pos = 0;
	const size_t n = text.size();
	std::string tag;
	auto nextToken = [&](std::string& out) -> bool {
		while (pos < n && isspace((unsigned char)text[pos]))
			pos++;
		if (pos >= n)
			return false;
		size_t j = pos;
		while (j < n && !isspace((unsigned char)text[j]))
			j++;
		out = text.substr(pos, j - pos);
		pos = j;
		return true;
	};
	auto restOfLine = [&]() -> std::string {
		size_t j = text.find('\n', pos);
		if (j == std::string::npos)
			j = n;
		std::string line = text.substr(pos, j - pos);
		pos = j < n ? j + 1 : n;
		return line;
	};
	if (!nextToken(tag) || tag != "solid")
		return NULL;
	Shared<TriMesh> obj = new TriMesh();
	int np = 0, indexv = 0, indexn = 0;
	while (nextToken(tag))
	{
		if (tag == "endfacet")
		{
			if (np == 3)
			{
				obj->indices << indexv - 3 << indexv - 2 << indexv - 1;
				obj->normalsI << indexn - 1 << indexn - 1 << indexn - 1;
			}
			np = 0;
		}
		else if (tag == "normal")
		{
			std::vector<float> a = floatsOf(restOfLine());
			a.resize(3, 0.0f);
			obj->normals << Vec3(a[0], a[1], a[2]);
			indexn++;
		}
		else if (tag == "vertex")
		{
			std::vector<float> a = floatsOf(restOfLine());
			a.resize(3, 0.0f);
			obj->vertices << Vec3(a[0], a[1], a[2]);
			np++;
			indexv++;
		}
	}
	return obj;
}

static Shared<TriMesh> loadSTLb(const std::string& data)
{
	// 80-byte header, int32 count, 50 bytes per facet: normal, 3 vertices, 2 attribute bytes (io.cpp:62-111)
	if (data.size() < 84)
		return NULL;
	int nf;
	memcpy(&nf, data.data() + 80, 4);
	if (nf < 0 || nf > 100000000)
		return NULL;
	Shared<TriMesh> obj = new TriMesh();
	obj->normals.reserve(nf);
	obj->vertices.reserve(nf * 3);
	obj->indices.reserve(nf * 3);
	obj->normalsI.reserve(nf * 3);
	int indexn = 0, indexv = 0;
	for (int i = 0; i < nf; i++, indexn++)
	{
		const size_t off = 84 + (size_t)i * 50;
		if (off + 50 > data.size())
			break;
		float f[12];
		memcpy(f, data.data() + off, 48);
		obj->normals << Vec3(f[0], f[1], f[2]);
		for (int j = 0; j < 3; j++, indexv++)
			obj->vertices << Vec3(f[3 + 3 * j], f[4 + 3 * j], f[5 + 3 * j]);
		obj->indices << indexv - 3 << indexv - 2 << indexv - 1;
		obj->normalsI << indexn << indexn << indexn;
	}
	return obj;
}

Shared<TriMesh> loadSTL(const String& filename)
{
	std::string data;
	if (!readFile(*filename, data) || data.size() < 5)
		return NULL;
	// binary iff the size matches the facet count at byte 80 (io.cpp:140-148)
	bool binary = false;
	if (data.size() > 84)
	{
		int nf;
		memcpy(&nf, data.data() + 80, 4);
		if ((long long)data.size() == 80LL + 4 + (long long)nf * 50)
			binary = true;
	}
	return binary ? loadSTLb(data) : loadSTLa(data);
}

void saveSTL(Shared<TriMesh> mesh, const String& name)
{
	// binary STL of the transformed triangles; facet normal = ((c - a) ^ (b - a)).normalized()
	// exactly as the reference writes it (io.cpp:158-185)
	FILE* f = fopen(*name, "wb");
	if (!f)
		return;
	char header[80];
	memset(header, ' ', sizeof(header));
	fwrite(header, 1, 80, f);
	const int nt = mesh->indices.length() / 3;
	fwrite(&nt, 4, 1, f);
	for (int i = 0; i + 2 < mesh->indices.length(); i += 3)
	{
		Vec3 a = mesh->vertices[mesh->indices[i]], b = mesh->vertices[mesh->indices[i + 1]], c = mesh->vertices[mesh->indices[i + 2]];
		a = mesh->transform * a;
		b = mesh->transform * b;
		c = mesh->transform * c;
		const Vec3 n = ((c - a) ^ (b - a)).normalized();
		const float rec[12] = { n.x, n.y, n.z, a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z };
		fwrite(rec, 4, 12, f);
		const short attr = 0;
		fwrite(&attr, 2, 1, f);
	}
	fclose(f);
}

void saveXYZ(const Array2<Vec3>& points, const String& filename, const Matrix4& m)
{
	// one "x y z" line per range-image point in front of the camera (io.cpp:417-431)
	FILE* f = fopen(*filename, "w");
	if (!f)
		return;
	for (int i = 0; i < points.rows(); i++)
		for (int j = 0; j < points.cols(); j++)
		{
			Vec3 p = points(i, j);
			if (p.z < -1e-5f)
			{
				p = m * p;
				fprintf(f, "%f %f %f\n", p.x, p.y, p.z);
			}
		}
	fclose(f);
}

// ------------------------------------------------------------------------------------------
// OBJ + MTL (reference src/io.cpp:189-335): one TriMesh per material, all sharing the file's
// vertex / normal / texcoord arrays; faces are polygons terminated by -1 and fan-triangulated.
// ------------------------------------------------------------------------------------------
Shared<SceneNode> loadOBJ(const String& filename)
{
	std::string text;
	if (!readFile(*filename, text))
		return NULL;
	const std::string dir = directoryOf(*filename);

	std::vector<std::string> order; // material names in order of first use
	std::map<std::string, Shared<TriMesh> > meshes;
	std::map<std::string, Shared<Material> > materials;
	std::vector<std::string> matOrder;
	materials[""] = new Material;
	matOrder.push_back("");
	Shared<TriMesh> mesh = new TriMesh();
	meshes[""] = mesh;
	order.push_back("");
	mesh->material = materials[""];

	Array<Vec3> vertices, normals;
	Array<Vec2> texcoords;

	size_t pos = 0;
	while (pos < text.size())
	{
		size_t eol = text.find('\n', pos);
		if (eol == std::string::npos)
			eol = text.size();
		std::string line = text.substr(pos, eol - pos);
		pos = eol + 1;
		if (!line.empty() && line[line.size() - 1] == '\r')
			line.erase(line.size() - 1);
		if (!line.empty() && line[0] == '#')
			continue;
		Strings parts = splitWs(line);
		if (parts.empty())
			continue;
		const std::string& key = parts[0];
		if (key == "v" && parts.size() >= 4)
			vertices << Vec3(toFloat(parts[1]), toFloat(parts[2]), toFloat(parts[3]));
		else if (key == "vn" && parts.size() >= 4)
			normals << Vec3(toFloat(parts[1]), toFloat(parts[2]), toFloat(parts[3]));
		else if (key == "vt" && parts.size() >= 3)
			texcoords << Vec2(toFloat(parts[1]), 1.0f - toFloat(parts[2])); // v flipped (io.cpp:233)
		else if (key == "f")
		{
			for (size_t i = 1; i < parts.size(); i++)
			{
				const Strings ix = splitChar(parts[i], '/');
				mesh->indices << toInt(ix[0]) - 1;
				if (ix.size() > 1)
					mesh->texcoordsI << toInt(ix[1]) - 1;
				if (ix.size() > 2)
					mesh->normalsI << toInt(ix[2]) - 1;
			}
			mesh->indices << -1;
			mesh->texcoordsI << -1;
			mesh->normalsI << -1;
		}
		else if (key == "usemtl" && parts.size() >= 2)
		{
			const std::string& name = parts[1];
			if (!meshes.count(name))
			{
				mesh = new TriMesh;
				// a material that has not been defined yet binds the default one (io.cpp:256)
				mesh->material = materials.count(name) ? materials[name] : materials[""];
				meshes[name] = mesh;
				order.push_back(name);
			}
			mesh = meshes[name];
		}
		else if (key == "mtllib" && parts.size() >= 2)
		{
			std::string mtl;
			if (!readFile(dir + "/" + parts[1], mtl))
				continue;
			Shared<Material> mat = materials[""];
			size_t mp = 0;
			while (mp < mtl.size())
			{
				size_t me = mtl.find('\n', mp);
				if (me == std::string::npos)
					me = mtl.size();
				const Strings mparts = splitWs(mtl.substr(mp, me - mp));
				mp = me + 1;
				if (mparts.empty())
					continue;
				const std::string& k = mparts[0];
				if (k == "newmtl" && mparts.size() >= 2)
				{
					mat = new Material;
					if (!materials.count(mparts[1]))
						matOrder.push_back(mparts[1]);
					materials[mparts[1]] = mat;
				}
				else if (k == "Kd" && mparts.size() >= 4)
					mat->diffuse = Vec3(toFloat(mparts[1]), toFloat(mparts[2]), toFloat(mparts[3]));
				else if (k == "Ks" && mparts.size() >= 4)
					mat->specular = Vec3(toFloat(mparts[1]), toFloat(mparts[2]), toFloat(mparts[3]));
				else if (k == "Ke" && mparts.size() >= 4)
					mat->emissive = Vec3(toFloat(mparts[1]), toFloat(mparts[2]), toFloat(mparts[3]));
				else if (k == "Ns" && mparts.size() >= 2)
				{
					mat->shininess = toFloat(mparts[1]);
					if (mat->shininess < 0.0001f)
						mat->shininess = 10;
				}
				else if (k == "d" && mparts.size() >= 2)
					mat->opacity = toFloat(mparts[1]);
				else if (k == "map_Kd" && mparts.size() >= 2)
					mat->textureName = String(mparts[1]);
			}
		}
	}

	for (size_t i = 0; i < matOrder.size(); i++)
	{
		Shared<Material>& mat = materials[matOrder[i]];
		if (mat->textureName.ok())
			mat->texture = loadPPM(String(dir + "/" + mat->textureName.std()));
	}

	Shared<SceneNode> node = new SceneNode;
	for (size_t k = 0; k < order.size(); k++)
	{
		Shared<TriMesh>& m = meshes[order[k]];
		m->indices = triangulateIndices(m->indices);
		m->texcoordsI = triangulateIndices(m->texcoordsI);
		m->normalsI = triangulateIndices(m->normalsI);
		m->vertices = vertices; // shared storage, like the reference (io.cpp:313-315)
		m->normals = normals;
		m->texcoords = texcoords;
		if (m->normals.length() == 0)
		{
			m->normals = Array<Vec3>(); // own array: the flat normals are per mesh (io.cpp:318 dup())
			flatNormals(m);
		}
		node->children << Shared<SceneNode>(m);
	}
	return node;
}

// ------------------------------------------------------------------------------------------
// X3D (reference src/x3d.cpp:35-203) on a minimal XML reader: elements, attributes, comments,
// declarations, self-closing tags, the five predefined entities.
// ------------------------------------------------------------------------------------------
namespace {

struct XmlNode
{
	std::string tag;
	std::vector<std::pair<std::string, std::string> > attrs;
	std::vector<XmlNode*> children;
	~XmlNode()
	{
		for (size_t i = 0; i < children.size(); i++)
			delete children[i];
	}
	bool has(const char* name) const
	{
		for (size_t i = 0; i < attrs.size(); i++)
			if (attrs[i].first == name)
				return true;
		return false;
	}
	std::string attr(const char* name, const char* def = "") const
	{
		// asl `e["a"] | "default"`: the default also replaces an empty value
		for (size_t i = 0; i < attrs.size(); i++)
			if (attrs[i].first == name)
				return attrs[i].second.empty() ? std::string(def) : attrs[i].second;
		return def;
	}
	const XmlNode* child(const char* t) const
	{
		for (size_t i = 0; i < children.size(); i++)
			if (children[i]->tag == t)
				return children[i];
		return 0;
	}
};

std::string decodeEntities(const std::string& s)
{
	std::string o;
	for (size_t i = 0; i < s.size(); i++)
	{
		if (s[i] == '&')
		{
			static const char* names[] = { "&quot;", "&apos;", "&amp;", "&lt;", "&gt;" };
			static const char chars[] = { '"', '\'', '&', '<', '>' };
			bool hit = false;
			for (int k = 0; k < 5; k++)
			{
				const size_t n = strlen(names[k]);
				if (s.compare(i, n, names[k]) == 0)
				{
					o.push_back(chars[k]);
					i += n - 1;
					hit = true;
					break;
				}
			}
			if (hit)
				continue;
		}
		o.push_back(s[i]);
	}
	return o;
}

struct XmlParser
{
	const std::string& s;
	size_t p;
	explicit XmlParser(const std::string& text) : s(text), p(0) {}

	void skipWs()
	{
		while (p < s.size() && isspace((unsigned char)s[p]))
			p++;
	}
	// skips text, comments, <? ?> and <! > up to the next element start; false at the end
	bool toElement()
	{
		for (;;)
		{
			const size_t lt = s.find('<', p);
			if (lt == std::string::npos)
				return false;
			p = lt;
			if (s.compare(p, 4, "<!--") == 0)
			{
				const size_t e = s.find("-->", p + 4);
				if (e == std::string::npos)
					return false;
				p = e + 3;
			}
			else if (s.compare(p, 2, "<?") == 0)
			{
				const size_t e = s.find("?>", p + 2);
				if (e == std::string::npos)
					return false;
				p = e + 2;
			}
			else if (s.compare(p, 2, "<!") == 0)
			{
				// DOCTYPE, possibly with an internal subset in [ ]
				int depth = 0;
				while (p < s.size())
				{
					if (s[p] == '[') depth++;
					else if (s[p] == ']') depth--;
					else if (s[p] == '>' && depth <= 0) { p++; break; }
					p++;
				}
			}
			else
				return true;
		}
	}
	XmlNode* parseElement()
	{
		// at '<' of an opening tag
		p++;
		size_t j = p;
		while (j < s.size() && !isspace((unsigned char)s[j]) && s[j] != '>' && s[j] != '/')
			j++;
		XmlNode* n = new XmlNode;
		n->tag = s.substr(p, j - p);
		p = j;
		for (;;)
		{
			skipWs();
			if (p >= s.size())
				return n;
			if (s[p] == '/')
			{
				const size_t e = s.find('>', p);
				p = e == std::string::npos ? s.size() : e + 1;
				return n; // self-closing
			}
			if (s[p] == '>')
			{
				p++;
				break;
			}
			size_t k = p;
			while (k < s.size() && s[k] != '=' && !isspace((unsigned char)s[k]) && s[k] != '>' && s[k] != '/')
				k++;
			const std::string name = s.substr(p, k - p);
			p = k;
			skipWs();
			std::string value;
			if (p < s.size() && s[p] == '=')
			{
				p++;
				skipWs();
				if (p < s.size() && (s[p] == '"' || s[p] == '\''))
				{
					const char q = s[p++];
					const size_t e = s.find(q, p);
					const size_t end = e == std::string::npos ? s.size() : e;
					value = decodeEntities(s.substr(p, end - p));
					p = end < s.size() ? end + 1 : end;
				}
			}
			if (!name.empty())
				n->attrs.push_back(std::make_pair(name, value));
			else if (p < s.size())
				p++;
		}
		// children until the matching close tag
		for (;;)
		{
			const size_t lt = s.find('<', p);
			if (lt == std::string::npos)
			{
				p = s.size();
				return n;
			}
			p = lt;
			if (s.compare(p, 2, "</") == 0)
			{
				const size_t e = s.find('>', p);
				p = e == std::string::npos ? s.size() : e + 1;
				return n;
			}
			if (!toElement())
				return n;
			if (s.compare(p, 2, "</") == 0)
				continue;
			n->children.push_back(parseElement());
		}
	}
	XmlNode* parseDocument()
	{
		if (!toElement())
			return 0;
		return parseElement();
	}
};

const XmlNode* findDef(const XmlNode* n, const std::string& name)
{
	if (n->attr("DEF") == name)
		return n;
	for (size_t i = 0; i < n->children.size(); i++)
		if (const XmlNode* r = findDef(n->children[i], name))
			return r;
	return 0;
}

Vec3 toVec3(const std::string& s)
{
	std::vector<float> a = floatsOf(s);
	a.resize(3, 0.0f);
	return Vec3(a[0], a[1], a[2]);
}

struct X3dReader
{
	std::string filename;
	XmlNode* doc;
	X3dReader() : doc(0) {}
	~X3dReader() { delete doc; }

	// USE="name" refers to the element with DEF="name" anywhere in the document (x3d.cpp:43-52)
	const XmlNode* get(const XmlNode* item) const
	{
		if (item && item->has("USE"))
			return findDef(doc, item->attr("USE"));
		return item;
	}

	Shared<SceneNode> sceneItem(const XmlNode* e)
	{
		if (e->tag == "Transform" || e->tag == "Group")
		{
			// T * R(axis, angle) * S (x3d.cpp:60-65)
			Strings rotation = splitWs(e->attr("rotation", "0 0 1 0"));
			Strings translation = splitWs(e->attr("translation", "0 0 0"));
			Strings scale = splitWs(e->attr("scale", "1 1 1"));
			rotation.resize(4, "0");
			translation.resize(3, "0");
			scale.resize(3, "1");
			Shared<SceneNode> node = new SceneNode;
			node->transform = Matrix4::translate(toFloat(translation[0]), toFloat(translation[1]), toFloat(translation[2])) *
			                  Matrix4::rotate(Vec3(toFloat(rotation[0]), toFloat(rotation[1]), toFloat(rotation[2])), toFloat(rotation[3])) *
			                  Matrix4::scale(Vec3(toFloat(scale[0]), toFloat(scale[1]), toFloat(scale[2])));
			for (size_t i = 0; i < e->children.size(); i++)
			{
				Shared<SceneNode> n = sceneItem(e->children[i]);
				if (n)
					node->children << n;
			}
			return node;
		}
		if (e->tag == "Shape")
			return shape(e);
		if (e->tag == "Inline")
		{
			std::string url = e->attr("url");
			std::string clean;
			for (size_t i = 0; i < url.size(); i++)
				if (url[i] != '"')
					clean.push_back(url[i]);
			X3dReader sub;
			return sub.load(directoryOf(filename) + "/" + clean);
		}
		return NULL;
	}

	Shared<SceneNode> shape(const XmlNode* e)
	{
		TriMesh* mesh = new TriMesh;
		Shared<SceneNode> holder(mesh);
		const XmlNode* appx = get(e->child("Appearance"));
		mesh->material = new Material;
		if (const XmlNode* mat = appx ? get(appx->child("Material")) : 0)
		{
			// defaults and the shininess * 8 rule: x3d.cpp:85-91
			mesh->material->diffuse = toVec3(mat->attr("diffuseColor", "0.7 0.75 0.8"));
			mesh->material->specular = toVec3(mat->attr("specularColor", "0.4 0.4 0.4"));
			mesh->material->emissive = toVec3(mat->attr("emissiveColor", "0 0 0"));
			mesh->material->shininess = toFloat(mat->attr("shininess", "0.5")) * 8;
		}
		if (const XmlNode* tex = appx ? get(appx->child("ImageTexture")) : 0)
		{
			// the url is used as written (x3d.cpp:93-94): a quoted MFString url keeps its quotes and finds no file, like the reference
			const std::string url = tex->attr("url");
			const std::string name = noExt(url) + ".ppm"; // textures are looked up as PPM (x3d.cpp:96)
			mesh->material->textureName = String(name);
			if (mesh->material->textureName.ok())
				mesh->material->texture = loadPPM(String(directoryOf(filename) + "/" + name));
		}
		const XmlNode* ifs = get(e->child("IndexedFaceSet"));
		const XmlNode* its = get(e->child("IndexedTriangleSet"));
		if (const XmlNode* g = ifs ? ifs : its)
		{
			const XmlNode* cn = get(g->child("Coordinate"));
			const XmlNode* nn = get(g->child("Normal"));
			const XmlNode* tn = get(g->child("TextureCoordinate"));
			const std::vector<float> verts = cn ? floatsOf(cn->attr("point")) : std::vector<float>();
			const std::vector<float> normals = nn ? floatsOf(nn->attr("vector")) : std::vector<float>();
			const std::vector<float> uvs = tn ? floatsOf(tn->attr("point")) : std::vector<float>();
			for (size_t i = 0; i + 2 < verts.size(); i += 3)
				mesh->vertices << Vec3(verts[i], verts[i + 1], verts[i + 2]);
			for (size_t i = 0; i + 2 < normals.size(); i += 3)
				mesh->normals << Vec3(normals[i], normals[i + 1], normals[i + 2]);
			for (size_t i = 0; i + 1 < uvs.size(); i += 2)
				mesh->texcoords << Vec2(uvs[i], 1 - uvs[i + 1]); // v flipped (x3d.cpp:119)
			if (ifs)
			{
				// polygons terminated by -1, fan-triangulated; a missing texCoordIndex / normalIndex
				// defaults to the (triangulated) coordIndex (x3d.cpp:121-131)
				const Array<int> ci = toArray(intsOf(g->attr("coordIndex")));
				const Array<int> ti = toArray(intsOf(g->attr("texCoordIndex")));
				const Array<int> ni = toArray(intsOf(g->attr("normalIndex")));
				mesh->indices = triangulateIndices(ci);
				mesh->texcoordsI = ti.length() == 0 ? mesh->indices.clone() : triangulateIndices(ti);
				mesh->normalsI = ni.length() == 0 ? mesh->indices.clone() : triangulateIndices(ni);
			}
			else
			{
				mesh->indices = toArray(intsOf(g->attr("index")));
				mesh->texcoordsI = mesh->indices.clone();
				mesh->normalsI = mesh->indices.clone();
			}
			if (mesh->normals.length() == 0)
				flatNormals(mesh);
			if (mesh->texcoords.length() == 0)
			{
				// a single dummy texcoord referenced by every corner (x3d.cpp:151-155)
				mesh->texcoords << Vec2(0, 0);
				mesh->texcoordsI = Array<int>(mesh->indices.length(), 0);
			}
		}
		return holder;
	}

	Shared<SceneNode> load(const std::string& name)
	{
		std::string text;
		if (!readFile(name, text))
			return NULL;
		XmlParser parser(text);
		doc = parser.parseDocument();
		if (!doc || doc->tag != "X3D")
			return NULL;
		const XmlNode* scene = doc->child("Scene");
		if (!scene)
			return NULL;
		filename = name;
		Shared<SceneNode> root = new SceneNode();
		for (size_t i = 0; i < scene->children.size(); i++)
		{
			Shared<SceneNode> node = sceneItem(scene->children[i]);
			if (node)
				root->children << node;
		}
		return root;
	}
};

}

Shared<SceneNode> loadX3D(const String& filename)
{
	X3dReader reader;
	return reader.load(*filename);
}

// Dispatch on the extension (reference src/io.cpp:113-132); unknown formats give an empty node.
Shared<SceneNode> loadMesh(const String& filename)
{
	const std::string name = *filename;
	if (hasExtension(name, "stl"))
	{
		Shared<TriMesh> mesh = loadSTL(filename);
		if (!mesh)
			return NULL;
		Shared<SceneNode> node = new SceneNode;
		node->children << Shared<SceneNode>(mesh);
		return node;
	}
	if (hasExtension(name, "obj"))
		return loadOBJ(filename);
	if (hasExtension(name, "x3d"))
		return loadX3D(filename);
	return new SceneNode;
}

}
