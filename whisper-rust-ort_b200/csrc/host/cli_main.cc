// whisper_b200_cli — drop-in for the reference's `whisper_ort_bench` binary (src/main.rs).
#include "../../../include/whisper_b200.h"
int main(int argc, char** argv) { return wb_cli_main(argc, (const char* const*)argv); }
