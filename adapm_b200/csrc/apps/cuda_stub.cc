// Stubs of the CUDA entry points for CPU-only builds of the core (scripts/sanitize.sh: ThreadSanitizer / AddressSanitizer).
#include "adapm/fabric.h"
#include "adapm/store.h"
namespace adapm {
namespace cudamem {
bool available() { return false; }
int device_count() { return 0; }
void set_device(int) {}
char* alloc_zeroed(uint64_t) { return nullptr; }
void free_dev(char*) {}
void export_handle(char*, unsigned char*) {}
char* import_handle(const unsigned char*) { return nullptr; }
void close_handle(char*) {}
void enable_peer(int, int) {}
bool vmm_supported(int, bool* mc) { if (mc) *mc = false; return false; }
uint64_t vmm_granularity(int, int, bool) { return 0; }
VmmHeap vmm_alloc(int, uint64_t) { return VmmHeap(); }
int vmm_export_fd(const VmmHeap&) { return -1; }
VmmHeap vmm_import_fd(int, int, uint64_t) { return VmmHeap(); }
void vmm_free(VmmHeap&) {}
unsigned long long mc_create(int, uint64_t, int*) { return 0; }
unsigned long long mc_import_fd(int) { return 0; }
void mc_add_device(unsigned long long, int) {}
char* mc_bind_and_map(unsigned long long, int, const VmmHeap&) { return nullptr; }
void mc_unmap(unsigned long long, int, char*, uint64_t) {}
}
std::unique_ptr<Backend> make_cuda_backend(const Options&, const Layout&, std::shared_ptr<Fabric>) { throw Error("no cuda in this build"); }
}
