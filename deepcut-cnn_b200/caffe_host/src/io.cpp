#include "caffe/util/io.hpp"

#include <cstdio>
#include <fstream>

namespace caffe {

bool ReadFileToString(const string& filename, string* out) {
  std::ifstream f(filename.c_str(), std::ios::in | std::ios::binary);
  if (!f) return false;
  f.seekg(0, std::ios::end);
  const std::streamoff n = f.tellg();
  f.seekg(0, std::ios::beg);
  out->resize(static_cast<size_t>(n));
  if (n > 0) f.read(&(*out)[0], n);
  return static_cast<bool>(f) || f.eof();
}

bool ReadProtoFromTextFile(const char* filename, pl::Message* proto) {
  string text;
  if (!ReadFileToString(filename, &text)) { LOG(ERROR) << "File not found: " << filename; return false; }
  string err;
  if (!proto->ParseFromTextString(text, &err)) { LOG(ERROR) << filename << ": " << err; return false; }
  return true;
}

void WriteProtoToTextFile(const pl::Message& proto, const char* filename) {
  std::ofstream f(filename);
  CHECK(f.good()) << "cannot open " << filename;
  f << proto.DebugString();
}

bool ReadProtoFromBinaryFile(const char* filename, pl::Message* proto) {
  string bytes;
  if (!ReadFileToString(filename, &bytes)) { LOG(ERROR) << "File not found: " << filename; return false; }
  return proto->ParseFromBinaryString(bytes);
}

void WriteProtoToBinaryFile(const pl::Message& proto, const char* filename) {
  std::ofstream f(filename, std::ios::out | std::ios::trunc | std::ios::binary);
  CHECK(f.good()) << "cannot open " << filename;
  const string bytes = proto.SerializeAsString();
  f.write(bytes.data(), static_cast<std::streamsize>(bytes.size()));
}

void ReadNetParamsFromTextFileOrDie(const string& param_file, NetParameter* param) {
  CHECK(ReadProtoFromTextFile(param_file, param)) << "Failed to parse NetParameter file: " << param_file;
  CHECK_GT(param->layer_size() + param->input_size(), 0)
      << param_file << " defines no `layer {}` entries; V0/V1 (`layers {}`) nets are not supported by deepcut-cnn_b200";
}

void ReadNetParamsFromBinaryFileOrDie(const string& param_file, NetParameter* param) {
  CHECK(ReadProtoFromBinaryFile(param_file, param)) << "Failed to parse NetParameter file: " << param_file;
}

}  // namespace caffe
