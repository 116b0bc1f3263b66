// srtb/io/npy.hpp — minimal NPY v1.0 writer (the reference saves the dynamic spectrum with the
// vendored cnpy: pipeline/write_signal_pipe.hpp:242-243, shape {batch_size, count}, complex64, C order;
// userspace/src/plot_spectrum.py:35-56 reads it back as [freq][time]).
#pragma once
#include <complex>
#include <cstdint>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace srtb {
namespace io {

template <typename T>
struct npy_descr;
template <> struct npy_descr<float> { static constexpr const char* value = "<f4"; };
template <> struct npy_descr<double> { static constexpr const char* value = "<f8"; };
template <> struct npy_descr<std::complex<float>> { static constexpr const char* value = "<c8"; };
template <> struct npy_descr<std::complex<double>> { static constexpr const char* value = "<c16"; };
template <> struct npy_descr<int8_t> { static constexpr const char* value = "|i1"; };
template <> struct npy_descr<uint8_t> { static constexpr const char* value = "|u1"; };

template <typename T>
inline void npy_save(const std::string& path, const T* data, const std::vector<size_t>& shape) {
  std::string dict = std::string("{'descr': '") + npy_descr<T>::value + "', 'fortran_order': False, 'shape': (";
  size_t total = 1;
  for (size_t i = 0; i < shape.size(); i++) {
    dict += std::to_string(shape[i]);
    if (shape.size() == 1 || i + 1 < shape.size()) dict += ", ";
    total *= shape[i];
  }
  dict += "), }";
  // header = magic(6) + version(2) + len(2) + dict, padded with spaces to a multiple of 64, ending in '\n'
  size_t unpadded = 10 + dict.size() + 1;
  const size_t pad = (64 - unpadded % 64) % 64;
  dict.append(pad, ' ');
  dict.push_back('\n');
  if (dict.size() > 65535) throw std::runtime_error("npy header too long");
  std::ofstream f(path, std::ios::binary | std::ios::trunc);
  if (!f) throw std::runtime_error("cannot open " + path);
  const char magic[8] = {'\x93', 'N', 'U', 'M', 'P', 'Y', 1, 0};
  f.write(magic, 8);
  const uint16_t len = static_cast<uint16_t>(dict.size());
  const char lenb[2] = {static_cast<char>(len & 0xff), static_cast<char>(len >> 8)};
  f.write(lenb, 2);
  f.write(dict.data(), (std::streamsize)dict.size());
  f.write(reinterpret_cast<const char*>(data), (std::streamsize)(total * sizeof(T)));
  if (!f) throw std::runtime_error("failed writing " + path);
}

}  // namespace io
}  // namespace srtb
