#ifndef tchem_b200_hpp
#define tchem_b200_hpp

// Extensions of the B200 drop-in that the reference has no counterpart for.
namespace tchem { namespace b200 {

// Number of GPUs one call with HOST tensors is spread over (contiguous slices of the batch, one host thread and one
// table per device, no collective). Default: the environment variable TCHEM_DEVICES
// ("all" or a count), else 1. Device tensors always run on their own device.
void set_host_devices(int n);
int host_devices();

} }

#endif
