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
}
std::unique_ptr<Backend> make_cuda_backend(const Options&, const Layout&, std::shared_ptr<Fabric>) { throw Error("no cuda in this build"); }
}
