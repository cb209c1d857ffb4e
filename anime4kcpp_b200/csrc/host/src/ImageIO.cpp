// ac::core::imdecode / imread / imwrite for the drop-in (reference: core/src/ImageIO.cpp:20-87, which delegates to the un-vendored
// stb_image / stb_image_write).  Own codecs on zlib, no third-party image library:
//   read : PNG (non-interlaced; gray / gray+alpha / RGB / RGBA / palette, 1-16 bit), BMP (uncompressed 8 / 24 / 32 bit), PNM (P5 / P6),
//          TGA (uncompressed gray / true colour)
//   write: .png (8-bit, zlib, per-row filter choice), .bmp, .tga -- the reference's extension dispatch; .jpg / .jpeg return false (no JPEG
//          codec here), as does any unknown extension
// Results are 8-bit images with `mode` channels (IMREAD_UNCHANGED keeps the file's), converted the way the reference's decoder does
// (luma = (77 r + 150 g + 29 b) >> 8, missing alpha = 255, 16-bit samples keep their high byte).  Failures give an empty image / false.
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <zlib.h>

#include "AC/Core/Image.hpp"

namespace
{
    using Bytes = std::vector<std::uint8_t>;
    struct Raw
    {
        int w = 0, h = 0, c = 0;
        Bytes px;       // tightly packed, 8 bit, c channels
        bool ok() const { return w > 0 && h > 0 && c >= 1 && c <= 4 && px.size() == static_cast<std::size_t>(w) * h * c; }
    };

    inline std::uint32_t be32(const std::uint8_t* p) { return (std::uint32_t{ p[0] } << 24) | (std::uint32_t{ p[1] } << 16) | (std::uint32_t{ p[2] } << 8) | p[3]; }
    inline std::uint32_t le32(const std::uint8_t* p) { return (std::uint32_t{ p[3] } << 24) | (std::uint32_t{ p[2] } << 16) | (std::uint32_t{ p[1] } << 8) | p[0]; }
    inline std::uint32_t le16(const std::uint8_t* p) { return (std::uint32_t{ p[1] } << 8) | p[0]; }
    inline std::uint8_t lumaOf(int r, int g, int b) { return static_cast<std::uint8_t>((r * 77 + g * 150 + b * 29) >> 8); }
    // images larger than this many samples are refused (the reference's Image holds its size in an int)
    constexpr std::uint64_t kMaxSamples = 0x7fffffffull;

    // ---- PNG ---------------------------------------------------------------------------------------------------------------------
    bool decodePng(const std::uint8_t* buf, std::size_t size, Raw& out)
    {
        static const std::uint8_t sig[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
        if (size < 8 + 25 || std::memcmp(buf, sig, 8) != 0) return false;
        std::size_t pos = 8;
        std::uint32_t w = 0, h = 0;
        int depth = 0, ctype = -1, interlace = 0;
        Bytes idat, plte, trns;
        bool end = false;
        while (!end && pos + 12 <= size)
        {
            const std::uint32_t len = be32(buf + pos);
            const std::uint8_t* type = buf + pos + 4;
            const std::uint8_t* data = buf + pos + 8;
            if (len > size - pos - 12) return false;
            if (!std::memcmp(type, "IHDR", 4))
            {
                if (len < 13) return false;
                w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
                if (data[10] != 0 || data[11] != 0) return false;
            }
            else if (!std::memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
            else if (!std::memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
            else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
            else if (!std::memcmp(type, "IEND", 4)) end = true;
            pos += 12 + static_cast<std::size_t>(len);
        }
        if (!w || !h || interlace != 0 || idat.empty()) return false;
        int nch;
        switch (ctype)
        {
        case 0: nch = 1; break;
        case 2: nch = 3; break;
        case 3: nch = 1; break;
        case 4: nch = 2; break;
        case 6: nch = 4; break;
        default: return false;
        }
        if (!(depth == 8 || depth == 16 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4)))) return false;
        if (ctype == 3 && (depth == 16 || plte.size() < 3)) return false;
        if (static_cast<std::uint64_t>(w) * h * 4 > kMaxSamples) return false;
        const std::size_t bpp = static_cast<std::size_t>((nch * depth + 7) / 8);                 // filter unit
        const std::size_t line = (static_cast<std::size_t>(w) * nch * depth + 7) / 8;
        Bytes raw((line + 1) * h);
        {
            uLongf have = static_cast<uLongf>(raw.size());
            if (uncompress(raw.data(), &have, idat.data(), static_cast<uLong>(idat.size())) != Z_OK || have != raw.size()) return false;
        }
        // unfilter in place (rows keep their leading filter byte)
        for (std::uint32_t y = 0; y < h; y++)
        {
            std::uint8_t* cur = raw.data() + (line + 1) * y + 1;
            const std::uint8_t* up = y ? cur - (line + 1) : nullptr;
            const int f = cur[-1];
            for (std::size_t i = 0; i < line; i++)
            {
                const int a = i >= bpp ? cur[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
                int pred;
                switch (f)
                {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4:
                {
                    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                    pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: return false;
                }
                cur[i] = static_cast<std::uint8_t>(cur[i] + pred);
            }
        }
        const bool pal = ctype == 3;
        const int outc = pal ? (trns.empty() ? 3 : 4) : ((ctype == 0 || ctype == 2) && !trns.empty() ? nch + 1 : nch);
        out.w = static_cast<int>(w); out.h = static_cast<int>(h); out.c = outc;
        out.px.resize(static_cast<std::size_t>(w) * h * outc);
        // colour-key transparency of gray / RGB images (tRNS holds one 16-bit sample per channel)
        int key[3] = { -1, -1, -1 };
        if (!pal && outc == nch + 1 && trns.size() >= static_cast<std::size_t>(2 * nch))
            for (int k = 0; k < nch; k++) key[k] = (trns[2 * k] << 8) | trns[2 * k + 1];
        for (std::uint32_t y = 0; y < h; y++)
        {
            const std::uint8_t* row = raw.data() + (line + 1) * y + 1;
            std::uint8_t* o = out.px.data() + static_cast<std::size_t>(y) * w * outc;
            for (std::uint32_t x = 0; x < w; x++)
            {
                int s[4] = { 0, 0, 0, 0 };        // samples at file depth
                if (depth == 8) for (int k = 0; k < nch; k++) s[k] = row[static_cast<std::size_t>(x) * nch + k];
                else if (depth == 16) for (int k = 0; k < nch; k++) s[k] = (row[(static_cast<std::size_t>(x) * nch + k) * 2] << 8) | row[(static_cast<std::size_t>(x) * nch + k) * 2 + 1];
                else
                {
                    const std::size_t bit = static_cast<std::size_t>(x) * depth;
                    s[0] = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
                }
                if (pal)
                {
                    const std::size_t i = static_cast<std::size_t>(s[0]);
                    const bool in = 3 * i + 2 < plte.size();
                    o[0] = in ? plte[3 * i] : 0; o[1] = in ? plte[3 * i + 1] : 0; o[2] = in ? plte[3 * i + 2] : 0;
                    if (outc == 4) o[3] = i < trns.size() ? trns[i] : 255;
                }
                else
                {
                    bool transparent = outc == nch + 1;
                    for (int k = 0; k < nch; k++)
                    {
                        if (outc == nch + 1 && s[k] != key[k]) transparent = false;
                        o[k] = depth == 16 ? static_cast<std::uint8_t>(s[k] >> 8)
                             : depth == 8 ? static_cast<std::uint8_t>(s[k])
                             : static_cast<std::uint8_t>(s[k] * 255 / ((1 << depth) - 1));
                    }
                    if (outc == nch + 1) o[nch] = transparent ? 0 : 255;
                }
                o += outc;
            }
        }
        return true;
    }

    // ---- BMP ---------------------------------------------------------------------------------------------------------------------
    bool decodeBmp(const std::uint8_t* buf, std::size_t size, Raw& out)
    {
        if (size < 54 || buf[0] != 'B' || buf[1] != 'M') return false;
        const std::uint32_t off = le32(buf + 10), hdr = le32(buf + 14);
        if (hdr < 40) return false;
        const std::int32_t w = static_cast<std::int32_t>(le32(buf + 18));
        std::int32_t h = static_cast<std::int32_t>(le32(buf + 22));
        const int bits = static_cast<int>(le16(buf + 28));
        const std::uint32_t comp = le32(buf + 30);
        if (w <= 0 || h == 0 || !(bits == 8 || bits == 24 || bits == 32) || !(comp == 0 || (comp == 3 && bits == 32))) return false;
        const bool flip = h > 0;
        if (h < 0) h = -h;
        if (static_cast<std::uint64_t>(w) * h * 4 > kMaxSamples) return false;
        const std::size_t line = (static_cast<std::size_t>(w) * bits / 8 + 3) & ~std::size_t{ 3 };
        if (off > size || line * h > size - off) return false;
        const std::uint8_t* pal = buf + 14 + hdr;
        const std::size_t npal = bits == 8 ? (le32(buf + 46) ? le32(buf + 46) : 256) : 0;
        if (bits == 8 && 14 + hdr + npal * 4 > size) return false;
        out.w = w; out.h = h; out.c = bits == 32 ? 4 : 3;
        out.px.resize(static_cast<std::size_t>(w) * h * out.c);
        for (int y = 0; y < h; y++)
        {
            const std::uint8_t* row = buf + off + line * static_cast<std::size_t>(flip ? h - 1 - y : y);
            std::uint8_t* o = out.px.data() + static_cast<std::size_t>(y) * w * out.c;
            for (int x = 0; x < w; x++, o += out.c)
            {
                if (bits == 8)
                {
                    const std::size_t i = row[x];
                    o[0] = i < npal ? pal[4 * i + 2] : 0; o[1] = i < npal ? pal[4 * i + 1] : 0; o[2] = i < npal ? pal[4 * i] : 0;
                }
                else
                {
                    const std::uint8_t* p = row + static_cast<std::size_t>(x) * (bits / 8);
                    o[0] = p[2]; o[1] = p[1]; o[2] = p[0];
                    if (bits == 32) o[3] = p[3];
                }
            }
        }
        return true;
    }

    // ---- PNM (binary gray / RGB) -------------------------------------------------------------------------------------------------
    bool decodePnm(const std::uint8_t* buf, std::size_t size, Raw& out)
    {
        if (size < 7 || buf[0] != 'P' || (buf[1] != '5' && buf[1] != '6')) return false;
        std::size_t pos = 2;
        int vals[3];
        for (int k = 0; k < 3; k++)
        {
            for (;;)
            {
                while (pos < size && std::isspace(buf[pos])) pos++;
                if (pos < size && buf[pos] == '#') { while (pos < size && buf[pos] != '\n') pos++; continue; }
                break;
            }
            long v = 0;
            int digits = 0;
            while (pos < size && std::isdigit(buf[pos]) && digits < 9) { v = v * 10 + (buf[pos] - '0'); pos++; digits++; }
            if (!digits) return false;
            vals[k] = static_cast<int>(v);
        }
        if (pos >= size || !std::isspace(buf[pos])) return false;
        pos++;
        const int c = buf[1] == '5' ? 1 : 3;
        if (vals[0] <= 0 || vals[1] <= 0 || vals[2] <= 0 || vals[2] > 255) return false;
        const std::uint64_t n = static_cast<std::uint64_t>(vals[0]) * vals[1] * c;
        if (n > kMaxSamples || n > size - pos) return false;
        out.w = vals[0]; out.h = vals[1]; out.c = c;
        out.px.assign(buf + pos, buf + pos + n);
        if (vals[2] != 255) for (auto& v : out.px) v = static_cast<std::uint8_t>(v * 255 / vals[2]);
        return true;
    }

    // ---- TGA (uncompressed) ------------------------------------------------------------------------------------------------------
    bool decodeTga(const std::uint8_t* buf, std::size_t size, Raw& out)
    {
        if (size < 18) return false;
        const int idlen = buf[0], cmap = buf[1], type = buf[2], w = static_cast<int>(le16(buf + 12)), h = static_cast<int>(le16(buf + 14)), bits = buf[16];
        if (cmap != 0 || !(type == 2 || type == 3) || w <= 0 || h <= 0) return false;
        if (!((type == 3 && bits == 8) || (type == 2 && (bits == 24 || bits == 32)))) return false;
        const int c = bits / 8;
        const std::size_t off = 18 + static_cast<std::size_t>(idlen), n = static_cast<std::size_t>(w) * h * c;
        if (off > size || n > size - off) return false;
        const bool top = (buf[17] & 0x20) != 0;
        out.w = w; out.h = h; out.c = c;
        out.px.resize(n);
        for (int y = 0; y < h; y++)
        {
            const std::uint8_t* row = buf + off + static_cast<std::size_t>(top ? y : h - 1 - y) * w * c;
            std::uint8_t* o = out.px.data() + static_cast<std::size_t>(y) * w * c;
            for (int x = 0; x < w; x++, row += c, o += c)
            {
                if (c == 1) o[0] = row[0];
                else { o[0] = row[2]; o[1] = row[1]; o[2] = row[0]; if (c == 4) o[3] = row[3]; }
            }
        }
        return true;
    }

    // the decoder's channel conversion: gray <-> colour, alpha dropped or filled with 255
    Raw convertChannels(const Raw& in, const int want)
    {
        if (want == in.c || want < 1 || want > 4) return in;
        Raw out;
        out.w = in.w; out.h = in.h; out.c = want;
        out.px.resize(static_cast<std::size_t>(in.w) * in.h * want);
        const std::uint8_t* s = in.px.data();
        std::uint8_t* d = out.px.data();
        for (std::size_t i = 0, n = static_cast<std::size_t>(in.w) * in.h; i < n; i++, s += in.c, d += want)
        {
            const int r = s[0], g = in.c >= 3 ? s[1] : s[0], b = in.c >= 3 ? s[2] : s[0];
            const int a = (in.c == 2) ? s[1] : (in.c == 4 ? s[3] : 255);
            const std::uint8_t y = in.c >= 3 ? lumaOf(r, g, b) : s[0];
            switch (want)
            {
            case 1: d[0] = y; break;
            case 2: d[0] = y; d[1] = static_cast<std::uint8_t>(a); break;
            case 3: d[0] = static_cast<std::uint8_t>(r); d[1] = static_cast<std::uint8_t>(g); d[2] = static_cast<std::uint8_t>(b); break;
            default: d[0] = static_cast<std::uint8_t>(r); d[1] = static_cast<std::uint8_t>(g); d[2] = static_cast<std::uint8_t>(b); d[3] = static_cast<std::uint8_t>(a); break;
            }
        }
        return out;
    }

    bool readFile(const char* filename, Bytes& out)
    {
        std::FILE* f = std::fopen(filename, "rb");
        if (!f) return false;
        bool ok = false;
        if (std::fseek(f, 0, SEEK_END) == 0)
        {
            const long n = std::ftell(f);
            if (n > 0 && std::fseek(f, 0, SEEK_SET) == 0)
            {
                out.resize(static_cast<std::size_t>(n));
                ok = std::fread(out.data(), 1, out.size(), f) == out.size();
            }
        }
        std::fclose(f);
        return ok;
    }
    bool writeFile(const char* filename, const Bytes& data)
    {
        std::FILE* f = std::fopen(filename, "wb");
        if (!f) return false;
        const bool ok = std::fwrite(data.data(), 1, data.size(), f) == data.size();
        return (std::fclose(f) == 0) && ok;
    }

    // ---- encoders ----------------------------------------------------------------------------------------------------------------
    void putBe32(Bytes& b, std::uint32_t v) { b.push_back(v >> 24); b.push_back((v >> 16) & 0xff); b.push_back((v >> 8) & 0xff); b.push_back(v & 0xff); }
    void pngChunk(Bytes& out, const char* type, const Bytes& data)
    {
        putBe32(out, static_cast<std::uint32_t>(data.size()));
        const std::size_t start = out.size();
        out.insert(out.end(), type, type + 4);
        out.insert(out.end(), data.begin(), data.end());
        putBe32(out, static_cast<std::uint32_t>(crc32(0L, out.data() + start, static_cast<uInt>(out.size() - start))));
    }
    bool encodePng(const ac::core::Image& img, Bytes& out)
    {
        const int w = img.width(), h = img.height(), c = img.channels();
        static const int ctypes[5] = { -1, 0, 4, 2, 6 };
        const std::size_t line = static_cast<std::size_t>(w) * c;
        // per row: the filter (none / sub / up / average / paeth) with the smallest sum of absolute residuals
        Bytes raw((line + 1) * h), trial(line);
        for (int y = 0; y < h; y++)
        {
            const std::uint8_t* cur = static_cast<const std::uint8_t*>(img.ptr()) + static_cast<std::ptrdiff_t>(y) * img.stride();
            const std::uint8_t* up = y ? cur - img.stride() : nullptr;
            std::uint8_t* dst = raw.data() + (line + 1) * y;
            long best = -1;
            for (int f = 0; f < 5; f++)
            {
                long cost = 0;
                for (std::size_t i = 0; i < line; i++)
                {
                    const int a = i >= static_cast<std::size_t>(c) ? cur[i - c] : 0, b = up ? up[i] : 0, cc = (up && i >= static_cast<std::size_t>(c)) ? up[i - c] : 0;
                    int pred = 0;
                    if (f == 1) pred = a;
                    else if (f == 2) pred = b;
                    else if (f == 3) pred = (a + b) >> 1;
                    else if (f == 4)
                    {
                        const int p = a + b - cc, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - cc);
                        pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : cc);
                    }
                    trial[i] = static_cast<std::uint8_t>(cur[i] - pred);
                    cost += std::abs(static_cast<int>(static_cast<std::int8_t>(trial[i])));
                }
                if (best < 0 || cost < best)
                {
                    best = cost;
                    dst[0] = static_cast<std::uint8_t>(f);
                    std::memcpy(dst + 1, trial.data(), line);
                }
            }
        }
        uLongf clen = compressBound(static_cast<uLong>(raw.size()));
        Bytes z(clen);
        if (compress2(z.data(), &clen, raw.data(), static_cast<uLong>(raw.size()), 6) != Z_OK) return false;
        z.resize(clen);
        static const std::uint8_t sig[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
        out.assign(sig, sig + 8);
        Bytes ihdr;
        putBe32(ihdr, static_cast<std::uint32_t>(w)); putBe32(ihdr, static_cast<std::uint32_t>(h));
        ihdr.push_back(8); ihdr.push_back(static_cast<std::uint8_t>(ctypes[c])); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
        pngChunk(out, "IHDR", ihdr);
        pngChunk(out, "IDAT", z);
        pngChunk(out, "IEND", Bytes{});
        return true;
    }
    void putLe16(Bytes& b, std::uint32_t v) { b.push_back(v & 0xff); b.push_back((v >> 8) & 0xff); }
    void putLe32(Bytes& b, std::uint32_t v) { putLe16(b, v & 0xffff); putLe16(b, v >> 16); }
    // colour samples of pixel x of a row as (r, g, b, a); gray is replicated
    inline void rgbaOf(const std::uint8_t* row, int x, int c, std::uint8_t (&p)[4])
    {
        const std::uint8_t* s = row + static_cast<std::ptrdiff_t>(x) * c;
        p[0] = s[0]; p[1] = c >= 3 ? s[1] : s[0]; p[2] = c >= 3 ? s[2] : s[0]; p[3] = (c == 2) ? s[1] : (c == 4 ? s[3] : 255);
    }
    bool encodeBmp(const ac::core::Image& img, Bytes& out)
    {
        const int w = img.width(), h = img.height(), c = img.channels();
        const int bytes = (c == 4 || c == 2) ? 4 : 3;
        const std::size_t line = (static_cast<std::size_t>(w) * bytes + 3) & ~std::size_t{ 3 };
        out.clear();
        out.push_back('B'); out.push_back('M');
        putLe32(out, static_cast<std::uint32_t>(54 + line * h)); putLe32(out, 0); putLe32(out, 54);
        putLe32(out, 40); putLe32(out, static_cast<std::uint32_t>(w)); putLe32(out, static_cast<std::uint32_t>(h)); putLe16(out, 1); putLe16(out, static_cast<std::uint32_t>(bytes * 8));
        putLe32(out, 0); putLe32(out, static_cast<std::uint32_t>(line * h)); putLe32(out, 2835); putLe32(out, 2835); putLe32(out, 0); putLe32(out, 0);
        for (int y = h - 1; y >= 0; y--)
        {
            const std::uint8_t* row = static_cast<const std::uint8_t*>(img.ptr()) + static_cast<std::ptrdiff_t>(y) * img.stride();
            const std::size_t start = out.size();
            for (int x = 0; x < w; x++)
            {
                std::uint8_t p[4];
                rgbaOf(row, x, c, p);
                out.push_back(p[2]); out.push_back(p[1]); out.push_back(p[0]);
                if (bytes == 4) out.push_back(p[3]);
            }
            out.resize(start + line, 0);
        }
        return true;
    }
    bool encodeTga(const ac::core::Image& img, Bytes& out)
    {
        const int w = img.width(), h = img.height(), c = img.channels();
        if (w > 0xffff || h > 0xffff) return false;
        const bool gray = c == 1;
        const int bytes = gray ? 1 : (c == 3 ? 3 : 4);
        out.assign(18, 0);
        out[2] = gray ? 3 : 2;
        out[12] = w & 0xff; out[13] = (w >> 8) & 0xff; out[14] = h & 0xff; out[15] = (h >> 8) & 0xff;
        out[16] = static_cast<std::uint8_t>(bytes * 8);
        out[17] = static_cast<std::uint8_t>(0x20 | (bytes == 4 ? 8 : 0));       // top-left origin, alpha bits
        for (int y = 0; y < h; y++)
        {
            const std::uint8_t* row = static_cast<const std::uint8_t*>(img.ptr()) + static_cast<std::ptrdiff_t>(y) * img.stride();
            for (int x = 0; x < w; x++)
            {
                std::uint8_t p[4];
                rgbaOf(row, x, c, p);
                if (gray) out.push_back(p[0]);
                else { out.push_back(p[2]); out.push_back(p[1]); out.push_back(p[0]); if (bytes == 4) out.push_back(p[3]); }
            }
        }
        return true;
    }
}

ac::core::Image ac::core::imdecode(const void* const buffer, const int size, const int mode) noexcept
{
    Image image{};
    if (!buffer || size <= 0) return image;
    try
    {
        const auto* p = static_cast<const std::uint8_t*>(buffer);
        const std::size_t n = static_cast<std::size_t>(size);
        Raw raw;
        if (!(decodePng(p, n, raw) || decodeBmp(p, n, raw) || decodePnm(p, n, raw) || decodeTga(p, n, raw)) || !raw.ok()) return image;
        if (mode > 0 && mode <= 4) raw = convertChannels(raw, mode);
        image.from(raw.w, raw.h, raw.c, Image::UInt8, raw.px.data());
    }
    catch (...) { image = Image{}; }
    return image;
}
ac::core::Image ac::core::imread(const char* const filename, const int mode) noexcept
{
    try
    {
        Bytes file;
        if (!filename || !readFile(filename, file) || file.size() > 0x7fffffffu) return Image{};
        return imdecode(file.data(), static_cast<int>(file.size()), mode);
    }
    catch (...) { return Image{}; }
}
bool ac::core::imwrite(const char* const filename, const Image& image) noexcept
{
    if (!filename || image.empty() || image.type() != Image::UInt8 || image.channels() < 1 || image.channels() > 4) return false;
    const char* point = std::strrchr(filename, '.');
    if (!point || !*(++point)) return false;
    char ext[5] = "";
    for (int i = 0; point[i] && i < 4; i++) ext[i] = static_cast<char>(std::tolower(static_cast<unsigned char>(point[i])));
    try
    {
        Bytes out;
        bool ok = false;
        if (!std::strcmp(ext, "png")) ok = encodePng(image, out);
        else if (!std::strcmp(ext, "bmp")) ok = encodeBmp(image, out);
        else if (!std::strcmp(ext, "tga")) ok = encodeTga(image, out);
        return ok && writeFile(filename, out);
    }
    catch (...) { return false; }
}
