// placeholder: dense per-geometry transforms (cart2int gradient, Hessians) - implemented next
#include "ic_table.h"
using namespace tchem_b200;
extern "C" {
int tchem_ic_grad_cart2int(const tchem_ic_table *, const double *, const double *, int64_t, double *, tchem_stream_t) {
    return set_error(TCHEM_EDOMAIN, "tchem_ic_grad_cart2int: not implemented yet");
}
int tchem_ic_grad_cart2int_matrix(const tchem_ic_table *, const double *, int64_t, double *, tchem_stream_t) {
    return set_error(TCHEM_EDOMAIN, "tchem_ic_grad_cart2int_matrix: not implemented yet");
}
int tchem_ic_hess_int2cart(const tchem_ic_table *, const double *, const double *, const double *, int64_t, double *, tchem_stream_t) {
    return set_error(TCHEM_EDOMAIN, "tchem_ic_hess_int2cart: not implemented yet");
}
int tchem_ic_hess_cart2int(const tchem_ic_table *, const double *, const double *, const double *, int64_t, double *, tchem_stream_t) {
    return set_error(TCHEM_EDOMAIN, "tchem_ic_hess_cart2int: not implemented yet");
}
int tchem_ic_sap_eval_host(const tchem_ic_table *, const tchem_sap_table *, const double *, int64_t, double *, double *, double *) {
    return set_error(TCHEM_EDOMAIN, "tchem_ic_sap_eval_host: not implemented yet");
}
}
