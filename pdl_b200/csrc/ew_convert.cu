// ew_convert.cu — device-side converttype (lib/PDL/Core/pdlconv.c:45-126): a(); [o]b()
// with b of another element type.  Keeps mixed-type expressions (float + 1.5,
// long * double) on the device instead of a CPU convert + PCIe round trip.
#include "elementwise.cuh"
#include "ew_ops.cuh"
namespace pdlb200 {
template <class TI>
static int convert_from(const pdlb200_trans *t, const Err &E) {
  switch (t->pdls[1].type) {
    case PDLB200_SB:  return ew_launch_typed<OpConvert, TI, int8_t,   1>(t, false, "ew_convert", E);
    case PDLB200_B:   return ew_launch_typed<OpConvert, TI, uint8_t,  1>(t, false, "ew_convert", E);
    case PDLB200_S:   return ew_launch_typed<OpConvert, TI, int16_t,  1>(t, false, "ew_convert", E);
    case PDLB200_US:  return ew_launch_typed<OpConvert, TI, uint16_t, 1>(t, false, "ew_convert", E);
    case PDLB200_L:   return ew_launch_typed<OpConvert, TI, int32_t,  1>(t, false, "ew_convert", E);
    case PDLB200_UL:  return ew_launch_typed<OpConvert, TI, uint32_t, 1>(t, false, "ew_convert", E);
    case PDLB200_IND: case PDLB200_LL: return ew_launch_typed<OpConvert, TI, int64_t, 1>(t, false, "ew_convert", E);
    case PDLB200_ULL: return ew_launch_typed<OpConvert, TI, uint64_t, 1>(t, false, "ew_convert", E);
    case PDLB200_F:   return ew_launch_typed<OpConvert, TI, float,    1>(t, false, "ew_convert", E);
    case PDLB200_D:   return ew_launch_typed<OpConvert, TI, double,   1>(t, false, "ew_convert", E);
    default: return E.fail(PDLB200_EUNSUPPORTED, "convert: target type %d is not on the device path", t->pdls[1].type);
  }
}
int launch_convert(const pdlb200_trans *t, const Err &E) {
  switch (t->datatype) {
    case PDLB200_SB:  return convert_from<int8_t>(t, E);
    case PDLB200_B:   return convert_from<uint8_t>(t, E);
    case PDLB200_S:   return convert_from<int16_t>(t, E);
    case PDLB200_US:  return convert_from<uint16_t>(t, E);
    case PDLB200_L:   return convert_from<int32_t>(t, E);
    case PDLB200_UL:  return convert_from<uint32_t>(t, E);
    case PDLB200_IND: case PDLB200_LL: return convert_from<int64_t>(t, E);
    case PDLB200_ULL: return convert_from<uint64_t>(t, E);
    case PDLB200_F:   return convert_from<float>(t, E);
    case PDLB200_D:   return convert_from<double>(t, E);
    default: return E.fail(PDLB200_EUNSUPPORTED, "convert: source type %d is not on the device path", t->datatype);
  }
}
}  // namespace pdlb200
