// Host build of the device CSV reader's number grammar (csrc/csv_number.h) for the CPU test-suite: the same source the kernels
// compile, driven over arrays of fields so that millions of literals can be checked against Python's float() / int().
#include "csv_number.h"
using namespace b200flow;

extern "C" {
// fields: concatenated bytes; offs[n + 1]; out_class / out_status / out_val / out_int per field
void csvnum_batch(const uint8_t* bytes, const int64_t* offs, int64_t n, int32_t* cls, int32_t* st, double* val, int32_t* st_i, int32_t* ival) {
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t* p = bytes + offs[i]; const int len = (int)(offs[i + 1] - offs[i]);
        cls[i] = csv_classify(p, len);
        double d = 0; st[i] = csv_parse_double(p, len, &d); val[i] = d;
        int32_t v = 0; st_i[i] = csv_parse_int32(p, len, &v); ival[i] = v;
    }
}
uint64_t csvnum_hash(const uint8_t* p, int len) { return csv_hash(p, len); }
}
