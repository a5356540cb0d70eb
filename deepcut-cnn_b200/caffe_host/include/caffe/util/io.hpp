// Proto file IO (reference include/caffe/util/io.hpp, src/caffe/util/io.cpp:34-77,
// src/caffe/util/upgrade_proto.cpp:66-78).  V0/V1 net upgrades are out of scope: the deepercut
// prototxt is V2 `layer {}` syntax; a file using `layers {}` fails with a clear message.
#pragma once
#include <string>

#include "caffe/common.hpp"
#include "caffe/proto/caffe.pb.h"

namespace caffe {

bool ReadFileToString(const string& filename, string* out);
bool ReadProtoFromTextFile(const char* filename, pl::Message* proto);
inline bool ReadProtoFromTextFile(const string& filename, pl::Message* proto) { return ReadProtoFromTextFile(filename.c_str(), proto); }
inline void ReadProtoFromTextFileOrDie(const string& filename, pl::Message* proto) { CHECK(ReadProtoFromTextFile(filename.c_str(), proto)) << "Failed to parse " << filename; }
void WriteProtoToTextFile(const pl::Message& proto, const char* filename);
bool ReadProtoFromBinaryFile(const char* filename, pl::Message* proto);
inline bool ReadProtoFromBinaryFile(const string& filename, pl::Message* proto) { return ReadProtoFromBinaryFile(filename.c_str(), proto); }
inline void ReadProtoFromBinaryFileOrDie(const string& filename, pl::Message* proto) { CHECK(ReadProtoFromBinaryFile(filename.c_str(), proto)) << "Failed to parse " << filename; }
void WriteProtoToBinaryFile(const pl::Message& proto, const char* filename);
inline void WriteProtoToBinaryFile(const pl::Message& proto, const string& filename) { WriteProtoToBinaryFile(proto, filename.c_str()); }
void ReadNetParamsFromTextFileOrDie(const string& param_file, NetParameter* param);
void ReadNetParamsFromBinaryFileOrDie(const string& param_file, NetParameter* param);

}  // namespace caffe
