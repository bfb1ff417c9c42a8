// PPM image I/O for the render path (reference src/io.cpp:337-415), on plain stdio.
#include <minirender/io.h>
#include <cstdio>
#include <vector>
#include <string>
#include <sstream>

using namespace asl;

namespace minirender {

void quantizeRGB8(const Array2<Vec3>& image, byte* out)
{
	for (int i = 0; i < image.rows(); i++)
		for (int j = 0; j < image.cols(); j++, out += 3)
		{
			const Vec3 v = image(i, j) * 255.0f;
			out[0] = (byte)clamp(v.x, 0.0f, 255.0f);
			out[1] = (byte)clamp(v.y, 0.0f, 255.0f);
			out[2] = (byte)clamp(v.z, 0.0f, 255.0f);
		}
}

void savePPM(const Array2<Vec3>& image, const String& filename)
{
	FILE* f = (filename == "--") ? stdout : fopen(*filename, "wb");
	if (!f)
	{
		printf("Cannot write file '%s'\n", *filename);
		return;
	}
	fprintf(f, "P6\n%d %d\n%d\n", image.cols(), image.rows(), 255);
	std::vector<byte> rgb((size_t)image.rows() * image.cols() * 3);
	if (!rgb.empty())
	{
		quantizeRGB8(image, rgb.data());
		fwrite(rgb.data(), 1, rgb.size(), f);
	}
	if (f != stdout)
		fclose(f);
}

Array2<Vec3> loadPPM(const String& filename)
{
	Array2<Vec3> image;
	FILE* f = fopen(*filename, "rb");
	if (!f)
		return image;
	// Header = everything up to the third newline that is not inside a '#' comment.
	std::string header;
	int newlines = 0;
	bool comment = false;
	int c;
	while (newlines < 3 && (c = fgetc(f)) != EOF)
	{
		if (c == '\n')
		{
			if (!comment)
				newlines++;
			comment = false;
		}
		else if (c == '#')
			comment = true;
		if (!comment)
			header.push_back((char)c);
	}
	std::istringstream in(header);
	std::vector<std::string> parts;
	for (std::string tok; in >> tok;)
		parts.push_back(tok);
	if (parts.size() != 4 || parts[0] != "P6")
	{
		fclose(f);
		return image;
	}
	const int cols = atoi(parts[1].c_str()), rows = atoi(parts[2].c_str());
	if (cols <= 0 || rows <= 0)
	{
		fclose(f);
		return image;
	}
	image.resize(rows, cols);
	for (int i = 0; i < rows; i++) // (rows a truncated file does not hold stay black; the reference leaves them unassigned)
		for (int j = 0; j < cols; j++)
			image(i, j) = Vec3(0, 0, 0);
	std::vector<byte> row((size_t)cols * 3);
	for (int i = 0; i < rows; i++)
	{
		if (fread(row.data(), 1, row.size(), f) < row.size())
			break;
		for (int j = 0; j < cols; j++)
		{
			const float r = row[3 * j], g = row[3 * j + 1], b = row[3 * j + 2];
			image(i, j) = Vec3(r, g, b) / 255.0f;
		}
	}
	fclose(f);
	return image;
}

}
