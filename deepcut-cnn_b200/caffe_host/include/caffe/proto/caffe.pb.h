// Look-alike of the protoc-generated caffe.pb.h for the messages the DeeperCut deploy path
// touches (field names, numbers and defaults from the reference's src/caffe/proto/caffe.proto:
// BlobShape :6, BlobProto :10, FillerParameter :43, NetParameter :64, NetState :258, NetStateRule
// :264, ParamSpec :288, LayerParameter :311, BatchNorm :501, Bias :513, Convolution :557,
// Crop :610, Eltwise :674, Pooling :855, ReLU :943, Scale :1014, Sigmoid :1038).
// Fields of caffe.proto that are not listed are skipped when parsing (text: with a warning).
#pragma once
#include "caffe/proto/proto_lite.hpp"

namespace caffe {

enum Phase { TRAIN = 0, TEST = 1 };
const pl::EnumTable& Phase_table();

enum FillerParameter_VarianceNorm { FillerParameter_VarianceNorm_FAN_IN = 0, FillerParameter_VarianceNorm_FAN_OUT = 1, FillerParameter_VarianceNorm_AVERAGE = 2 };
const pl::EnumTable& VarianceNorm_table();
enum ParamSpec_DimCheckMode { ParamSpec_DimCheckMode_STRICT = 0, ParamSpec_DimCheckMode_PERMISSIVE = 1 };
const pl::EnumTable& DimCheckMode_table();
enum Engine { Engine_DEFAULT = 0, Engine_CAFFE = 1, Engine_CUDNN = 2 };
const pl::EnumTable& Engine_table();
typedef Engine ConvolutionParameter_Engine;
typedef Engine PoolingParameter_Engine;
typedef Engine ReLUParameter_Engine;
typedef Engine SigmoidParameter_Engine;
enum EltwiseParameter_EltwiseOp { EltwiseParameter_EltwiseOp_PROD = 0, EltwiseParameter_EltwiseOp_SUM = 1, EltwiseParameter_EltwiseOp_MAX = 2 };
const pl::EnumTable& EltwiseOp_table();
enum PoolingParameter_PoolMethod { PoolingParameter_PoolMethod_MAX = 0, PoolingParameter_PoolMethod_AVE = 1, PoolingParameter_PoolMethod_STOCHASTIC = 2 };
const pl::EnumTable& PoolMethod_table();

#define BLOBSHAPE_FIELDS(OPT, REP, MSG, RMSG, ENM) REP(int64_t, dim, 1, true)
PL_DECLARE_MESSAGE(BlobShape, BLOBSHAPE_FIELDS)

#define BLOBPROTO_FIELDS(OPT, REP, MSG, RMSG, ENM) \
  MSG(BlobShape, shape, 7)                         \
  REP(float, data, 5, true)                        \
  REP(float, diff, 6, true)                        \
  REP(double, double_data, 8, true)                \
  REP(double, double_diff, 9, true)                \
  OPT(int32_t, num, 1, 0)                          \
  OPT(int32_t, channels, 2, 0)                     \
  OPT(int32_t, height, 3, 0)                       \
  OPT(int32_t, width, 4, 0)
PL_DECLARE_MESSAGE(BlobProto, BLOBPROTO_FIELDS)

#define FILLER_FIELDS(OPT, REP, MSG, RMSG, ENM)    \
  OPT(std::string, type, 1, "constant")            \
  OPT(float, value, 2, 0.f)                        \
  OPT(float, min, 3, 0.f)                          \
  OPT(float, max, 4, 1.f)                          \
  OPT(float, mean, 5, 0.f)                         \
  OPT(float, std, 6, 1.f)                          \
  OPT(int32_t, sparse, 7, -1)                      \
  ENM(FillerParameter_VarianceNorm, variance_norm, 8, FillerParameter_VarianceNorm_FAN_IN, VarianceNorm_table)
PL_DECLARE_MESSAGE(FillerParameter, FILLER_FIELDS)

#define NETSTATE_FIELDS(OPT, REP, MSG, RMSG, ENM)  \
  ENM(Phase, phase, 1, TEST, Phase_table)          \
  OPT(int32_t, level, 2, 0)                        \
  REP(std::string, stage, 3, false)
PL_DECLARE_MESSAGE(NetState, NETSTATE_FIELDS)

#define NETSTATERULE_FIELDS(OPT, REP, MSG, RMSG, ENM) \
  ENM(Phase, phase, 1, TRAIN, Phase_table)            \
  OPT(int32_t, min_level, 2, 0)                       \
  OPT(int32_t, max_level, 3, 0)                       \
  REP(std::string, stage, 4, false)                   \
  REP(std::string, not_stage, 5, false)
PL_DECLARE_MESSAGE(NetStateRule, NETSTATERULE_FIELDS)

#define PARAMSPEC_FIELDS(OPT, REP, MSG, RMSG, ENM)    \
  OPT(std::string, name, 1, "")                       \
  ENM(ParamSpec_DimCheckMode, share_mode, 2, ParamSpec_DimCheckMode_STRICT, DimCheckMode_table) \
  OPT(float, lr_mult, 3, 1.f)                         \
  OPT(float, decay_mult, 4, 1.f)
PL_DECLARE_MESSAGE(ParamSpec, PARAMSPEC_FIELDS)

#define BATCHNORM_FIELDS(OPT, REP, MSG, RMSG, ENM)    \
  OPT(bool, use_global_stats, 1, false)               \
  OPT(float, moving_average_fraction, 2, .999f)       \
  OPT(float, eps, 3, 1e-5f)
PL_DECLARE_MESSAGE(BatchNormParameter, BATCHNORM_FIELDS)

#define BIAS_FIELDS(OPT, REP, MSG, RMSG, ENM)         \
  OPT(int32_t, axis, 1, 1)                            \
  OPT(int32_t, num_axes, 2, 1)                        \
  MSG(FillerParameter, filler, 3)
PL_DECLARE_MESSAGE(BiasParameter, BIAS_FIELDS)

#define CONVOLUTION_FIELDS(OPT, REP, MSG, RMSG, ENM)  \
  OPT(uint32_t, num_output, 1, 0)                     \
  OPT(bool, bias_term, 2, true)                       \
  REP(uint32_t, pad, 3, false)                        \
  REP(uint32_t, kernel_size, 4, false)                \
  REP(uint32_t, stride, 6, false)                     \
  REP(uint32_t, dilation, 18, false)                  \
  OPT(uint32_t, pad_h, 9, 0)                          \
  OPT(uint32_t, pad_w, 10, 0)                         \
  OPT(uint32_t, kernel_h, 11, 0)                      \
  OPT(uint32_t, kernel_w, 12, 0)                      \
  OPT(uint32_t, stride_h, 13, 0)                      \
  OPT(uint32_t, stride_w, 14, 0)                      \
  OPT(uint32_t, group, 5, 1)                          \
  MSG(FillerParameter, weight_filler, 7)              \
  MSG(FillerParameter, bias_filler, 8)                \
  ENM(Engine, engine, 15, Engine_DEFAULT, Engine_table) \
  OPT(int32_t, axis, 16, 1)                           \
  OPT(bool, force_nd_im2col, 17, false)
PL_DECLARE_MESSAGE(ConvolutionParameter, CONVOLUTION_FIELDS)

#define CROP_FIELDS(OPT, REP, MSG, RMSG, ENM)         \
  OPT(uint32_t, offset_height, 1, 0)                  \
  OPT(uint32_t, offset_width, 2, 0)
PL_DECLARE_MESSAGE(CropParameter, CROP_FIELDS)

#define ELTWISE_FIELDS(OPT, REP, MSG, RMSG, ENM)      \
  ENM(EltwiseParameter_EltwiseOp, operation, 1, EltwiseParameter_EltwiseOp_SUM, EltwiseOp_table) \
  REP(float, coeff, 2, false)                         \
  OPT(bool, stable_prod_grad, 3, true)
PL_DECLARE_MESSAGE(EltwiseParameter, ELTWISE_FIELDS)

#define POOLING_FIELDS(OPT, REP, MSG, RMSG, ENM)      \
  ENM(PoolingParameter_PoolMethod, pool, 1, PoolingParameter_PoolMethod_MAX, PoolMethod_table) \
  OPT(uint32_t, pad, 4, 0)                            \
  OPT(uint32_t, pad_h, 9, 0)                          \
  OPT(uint32_t, pad_w, 10, 0)                         \
  OPT(uint32_t, kernel_size, 2, 0)                    \
  OPT(uint32_t, kernel_h, 5, 0)                       \
  OPT(uint32_t, kernel_w, 6, 0)                       \
  OPT(uint32_t, stride, 3, 1)                         \
  OPT(uint32_t, stride_h, 7, 0)                       \
  OPT(uint32_t, stride_w, 8, 0)                       \
  ENM(Engine, engine, 11, Engine_DEFAULT, Engine_table) \
  OPT(bool, global_pooling, 12, false)
PL_DECLARE_MESSAGE(PoolingParameter, POOLING_FIELDS)

#define RELU_FIELDS(OPT, REP, MSG, RMSG, ENM)         \
  OPT(float, negative_slope, 1, 0.f)                  \
  ENM(Engine, engine, 2, Engine_DEFAULT, Engine_table)
PL_DECLARE_MESSAGE(ReLUParameter, RELU_FIELDS)

#define SCALE_FIELDS(OPT, REP, MSG, RMSG, ENM)        \
  OPT(int32_t, axis, 1, 1)                            \
  OPT(int32_t, num_axes, 2, 1)                        \
  MSG(FillerParameter, filler, 3)                     \
  OPT(bool, bias_term, 4, false)                      \
  MSG(FillerParameter, bias_filler, 5)
PL_DECLARE_MESSAGE(ScaleParameter, SCALE_FIELDS)

#define SIGMOID_FIELDS(OPT, REP, MSG, RMSG, ENM) ENM(Engine, engine, 1, Engine_DEFAULT, Engine_table)
PL_DECLARE_MESSAGE(SigmoidParameter, SIGMOID_FIELDS)

#define LAYER_FIELDS(OPT, REP, MSG, RMSG, ENM)        \
  OPT(std::string, name, 1, "")                       \
  OPT(std::string, type, 2, "")                       \
  REP(std::string, bottom, 3, false)                  \
  REP(std::string, top, 4, false)                     \
  ENM(Phase, phase, 10, TEST, Phase_table)            \
  REP(float, loss_weight, 5, false)                   \
  RMSG(ParamSpec, param, 6)                           \
  RMSG(BlobProto, blobs, 7)                           \
  RMSG(NetStateRule, include, 8)                      \
  RMSG(NetStateRule, exclude, 9)                      \
  MSG(BatchNormParameter, batch_norm_param, 139)      \
  MSG(BiasParameter, bias_param, 141)                 \
  MSG(ConvolutionParameter, convolution_param, 106)   \
  MSG(CropParameter, crop_param, 143)                 \
  MSG(EltwiseParameter, eltwise_param, 110)           \
  MSG(PoolingParameter, pooling_param, 121)           \
  MSG(ReLUParameter, relu_param, 123)                 \
  MSG(ScaleParameter, scale_param, 142)               \
  MSG(SigmoidParameter, sigmoid_param, 124)
PL_DECLARE_MESSAGE(LayerParameter, LAYER_FIELDS)

#define NET_FIELDS(OPT, REP, MSG, RMSG, ENM)          \
  OPT(std::string, name, 1, "")                       \
  REP(std::string, input, 3, false)                   \
  RMSG(BlobShape, input_shape, 8)                     \
  REP(int32_t, input_dim, 4, false)                   \
  OPT(bool, force_backward, 5, false)                 \
  MSG(NetState, state, 6)                             \
  OPT(bool, debug_info, 7, false)                     \
  RMSG(LayerParameter, layer, 100)
PL_DECLARE_MESSAGE(NetParameter, NET_FIELDS)

}  // namespace caffe
