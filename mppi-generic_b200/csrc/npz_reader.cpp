/*
 * npz_reader.cpp — reads one array out of a NumPy .npz archive, the file format of the reference's model / map inputs
 * (cnpy::npz_load in FNNHelper::loadParams, utils/nn_helpers/fnn_helper.cu:44-56, and ARStandardCostImpl::loadTrackData,
 * cost_functions/autorally/ar_standard_cost.cu:85-142). cnpy is a git submodule of the reference (submodules/cnpy) and is
 * not vendored, so the two published formats are read directly: ZIP (PKWARE APPNOTE: end-of-central-directory record,
 * central directory, local headers, zip64 extra fields; methods 0 = stored and 8 = deflate through zlib) and NPY
 * (numpy.lib.format versions 1-3: magic, header length, a Python dict literal with descr / fortran_order / shape).
 * Host-only; nothing here touches the GPU.
 */
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "../../include/mppi_b200/host_twins.h"

extern "C" int mppib_set_last_error(int status, const char* fmt, ...);  // engine.cu

namespace
{
struct File
{
  FILE* f = nullptr;
  ~File()
  {
    if (f)
      fclose(f);
  }
};
inline uint16_t rd16(const unsigned char* p)
{
  return (uint16_t)(p[0] | (p[1] << 8));
}
inline uint32_t rd32(const unsigned char* p)
{
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
inline uint64_t rd64(const unsigned char* p)
{
  return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32);
}
bool read_at(FILE* f, uint64_t off, void* dst, size_t n)
{
  if (fseeko(f, (off_t)off, SEEK_SET) != 0)
    return false;
  return fread(dst, 1, n, f) == n;
}

struct Member
{
  uint16_t method = 0;
  uint64_t csize = 0, usize = 0, local_off = 0;
};

// locate `member` (e.g. "dynamics_W1.npy") through the central directory
int find_member(FILE* f, const std::string& member, Member& out)
{
  if (fseeko(f, 0, SEEK_END) != 0)
    return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: cannot seek");
  const uint64_t fsize = (uint64_t)ftello(f);
  const uint64_t tail = fsize < 66000 ? fsize : 66000;  // EOCD (22 B) + comment (<= 65535 B)
  std::vector<unsigned char> buf(tail);
  if (!read_at(f, fsize - tail, buf.data(), tail))
    return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: cannot read the end of the archive");
  long eocd = -1;
  for (long i = (long)tail - 22; i >= 0; i--)
    if (rd32(&buf[i]) == 0x06054b50u)
    {
      eocd = i;
      break;
    }
  if (eocd < 0)
    return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: not a zip archive (no end-of-central-directory record)");
  uint64_t entries = rd16(&buf[eocd + 10]), cd_size = rd32(&buf[eocd + 12]), cd_off = rd32(&buf[eocd + 16]);
  if (entries == 0xFFFFu || cd_size == 0xFFFFFFFFu || cd_off == 0xFFFFFFFFu)
  {  // zip64: locator sits right before the EOCD
    if (eocd < 20 || rd32(&buf[eocd - 20]) != 0x07064b50u)
      return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: zip64 locator missing");
    const uint64_t e64 = rd64(&buf[eocd - 20 + 8]);
    unsigned char r[56];
    if (!read_at(f, e64, r, 56) || rd32(r) != 0x06064b50u)
      return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: zip64 end-of-central-directory record unreadable");
    entries = rd64(r + 32);
    cd_size = rd64(r + 40);
    cd_off = rd64(r + 48);
  }
  if (cd_size > fsize || cd_off > fsize)
    return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: central directory outside the file");
  std::vector<unsigned char> cd(cd_size);
  if (cd_size == 0 || !read_at(f, cd_off, cd.data(), cd_size))
    return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: central directory unreadable");
  size_t p = 0;
  for (uint64_t e = 0; e < entries && p + 46 <= cd.size(); e++)
  {
    if (rd32(&cd[p]) != 0x02014b50u)
      return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: corrupt central directory");
    const uint16_t method = rd16(&cd[p + 10]);
    uint64_t csize = rd32(&cd[p + 20]), usize = rd32(&cd[p + 24]), loff = rd32(&cd[p + 42]);
    const uint16_t nlen = rd16(&cd[p + 28]), xlen = rd16(&cd[p + 30]), clen = rd16(&cd[p + 32]);
    if (p + 46 + nlen + xlen + clen > cd.size())
      return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: corrupt central directory entry");
    const std::string name((const char*)&cd[p + 46], nlen);
    // zip64 extended information (header id 1): the 0xFFFFFFFF fields, in this order
    size_t x = p + 46 + nlen, xend = x + xlen;
    while (x + 4 <= xend)
    {
      const uint16_t id = rd16(&cd[x]), sz = rd16(&cd[x + 2]);
      if (id == 0x0001)
      {
        size_t q = x + 4;
        if (usize == 0xFFFFFFFFu && q + 8 <= xend)
          usize = rd64(&cd[q]), q += 8;
        if (csize == 0xFFFFFFFFu && q + 8 <= xend)
          csize = rd64(&cd[q]), q += 8;
        if (loff == 0xFFFFFFFFu && q + 8 <= xend)
          loff = rd64(&cd[q]), q += 8;
      }
      x += 4 + sz;
    }
    if (name == member)
    {
      out.method = method;
      out.csize = csize;
      out.usize = usize;
      out.local_off = loff;
      return MPPIB_OK;
    }
    p += 46 + nlen + xlen + clen;
  }
  return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: no array named '%s'", member.c_str());
}

int read_member(FILE* f, const Member& m, std::vector<unsigned char>& data)
{
  unsigned char lh[30];
  if (!read_at(f, m.local_off, lh, 30) || rd32(lh) != 0x04034b50u)
    return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: local file header unreadable");
  const uint64_t data_off = m.local_off + 30 + rd16(lh + 26) + rd16(lh + 28);
  if (fseeko(f, 0, SEEK_END) != 0)
    return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: cannot seek");
  const uint64_t fsize = (uint64_t)ftello(f);
  if (m.csize > fsize || data_off > fsize || m.usize > (1ull << 34))
    return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: member sizes do not fit the file");
  std::vector<unsigned char> comp(m.csize);
  if (m.csize && !read_at(f, data_off, comp.data(), m.csize))
    return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: member data truncated");
  if (m.method == 0)
  {
    data.swap(comp);
    return MPPIB_OK;
  }
  if (m.method != 8)
    return mppib_set_last_error(MPPIB_ERR_UNSUPPORTED, "npz: compression method %d (only stored and deflate)", (int)m.method);
  data.resize(m.usize);
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  if (inflateInit2(&zs, -MAX_WBITS) != Z_OK)  // raw deflate stream, as zip stores it
    return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: inflateInit2 failed");
  zs.next_in = comp.data();
  zs.avail_in = (uInt)comp.size();
  zs.next_out = data.data();
  zs.avail_out = (uInt)data.size();
  const int zr = inflate(&zs, Z_FINISH);
  inflateEnd(&zs);
  if (zr != Z_STREAM_END || zs.total_out != m.usize)
    return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: inflate failed (%d)", zr);
  return MPPIB_OK;
}

// value of key in the header dict, e.g. "'descr': '<f8'" -> "<f8" ; "'shape': (32, 6)" -> "32, 6"
bool dict_value(const std::string& hdr, const char* key, char open, char close, std::string& out)
{
  const size_t k = hdr.find(std::string("'") + key + "'");
  if (k == std::string::npos)
    return false;
  const size_t colon = hdr.find(':', k);
  if (colon == std::string::npos)
    return false;
  if (open == 0)
  {  // bare word (True / False)
    size_t b = colon + 1;
    while (b < hdr.size() && hdr[b] == ' ')
      b++;
    size_t e = b;
    while (e < hdr.size() && hdr[e] != ',' && hdr[e] != '}' && hdr[e] != ' ')
      e++;
    out = hdr.substr(b, e - b);
    return true;
  }
  const size_t b = hdr.find(open, colon);
  if (b == std::string::npos)
    return false;
  const size_t e = hdr.find(close, b + 1);
  if (e == std::string::npos)
    return false;
  out = hdr.substr(b + 1, e - b - 1);
  return true;
}
}  // namespace

static int npz_read_impl(const char* path, const char* name, float* out, size_t capacity, size_t* count, int* shape4,
                         int* ndim);

extern "C" int mppib_host_npz_read(const char* path, const char* name, float* out, size_t capacity, size_t* count,
                                   int* shape4, int* ndim)
{
  try
  {  // no exception may cross the C ABI (a corrupt archive can ask for absurd buffer sizes)
    return npz_read_impl(path, name, out, capacity, count, shape4, ndim);
  }
  catch (const std::exception& e)
  {
    return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: %s", e.what());
  }
}

static int npz_read_impl(const char* path, const char* name, float* out, size_t capacity, size_t* count, int* shape4,
                         int* ndim)
{
  if (!path || !name)
    return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: null argument");
  File file;
  file.f = fopen(path, "rb");
  if (!file.f)
    return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: cannot open '%s'", path);
  Member m;
  int rc = find_member(file.f, std::string(name) + ".npy", m);
  if (rc != MPPIB_OK)
    return rc;
  std::vector<unsigned char> d;
  rc = read_member(file.f, m, d);
  if (rc != MPPIB_OK)
    return rc;
  if (d.size() < 12 || memcmp(d.data(), "\x93NUMPY", 6) != 0)
    return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: '%s' is not an NPY array", name);
  const int major = d[6];
  size_t hlen, hoff;
  if (major == 1)
    hlen = rd16(&d[8]), hoff = 10;
  else if (major == 2 || major == 3)
    hlen = rd32(&d[8]), hoff = 12;
  else
    return mppib_set_last_error(MPPIB_ERR_UNSUPPORTED, "npz: NPY format version %d", major);
  if (hoff + hlen > d.size())
    return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: truncated NPY header");
  const std::string hdr((const char*)&d[hoff], hlen);
  std::string descr, order, shape;
  if (!dict_value(hdr, "descr", '\'', '\'', descr) || !dict_value(hdr, "fortran_order", 0, 0, order) ||
      !dict_value(hdr, "shape", '(', ')', shape))
    return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: cannot parse the NPY header of '%s'", name);
  if (order != "False")
    return mppib_set_last_error(MPPIB_ERR_UNSUPPORTED, "npz: '%s' is Fortran-ordered", name);
  int dims[4] = { 1, 1, 1, 1 }, nd = 0;
  size_t n = 1;
  {
    size_t i = 0;
    while (i < shape.size())
    {
      while (i < shape.size() && (shape[i] == ' ' || shape[i] == ','))
        i++;
      if (i >= shape.size())
        break;
      size_t j = i;
      long v = 0;
      while (j < shape.size() && shape[j] >= '0' && shape[j] <= '9')
        v = v * 10 + (shape[j] - '0'), j++;
      if (j == i)
        return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: bad shape '%s'", shape.c_str());
      if (nd >= 4)
        return mppib_set_last_error(MPPIB_ERR_UNSUPPORTED, "npz: more than 4 dimensions");
      dims[nd++] = (int)v;
      n *= (size_t)v;
      i = j;
    }
  }
  size_t esize = 0;
  if (descr == "<f4" || descr == "<i4")
    esize = 4;
  else if (descr == "<f8" || descr == "<i8")
    esize = 8;
  else
    return mppib_set_last_error(MPPIB_ERR_UNSUPPORTED, "npz: dtype '%s' (float32 / float64 / int32 / int64, little-endian)",
                                descr.c_str());
  if (hoff + hlen + n * esize > d.size())
    return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: '%s' holds fewer bytes than its shape says", name);
  if (count)
    *count = n;
  if (ndim)
    *ndim = nd;
  if (shape4)
    for (int i = 0; i < 4; i++)
      shape4[i] = dims[i];
  if (!out)
    return MPPIB_OK;
  if (capacity < n)
    return mppib_set_last_error(MPPIB_ERR_INVALID_ARG, "npz: '%s' has %zu elements, the buffer holds %zu", name, n, capacity);
  const unsigned char* src = &d[hoff + hlen];
  for (size_t i = 0; i < n; i++)
  {
    if (descr == "<f4")
    {
      float v;
      memcpy(&v, src + 4 * i, 4);
      out[i] = v;
    }
    else if (descr == "<f8")
    {
      double v;
      memcpy(&v, src + 8 * i, 8);
      out[i] = (float)v;
    }
    else if (descr == "<i4")
    {
      int32_t v;
      memcpy(&v, src + 4 * i, 4);
      out[i] = (float)v;
    }
    else
    {
      int64_t v;
      memcpy(&v, src + 8 * i, 8);
      out[i] = (float)v;
    }
  }
  return MPPIB_OK;
}
