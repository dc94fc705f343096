/* LD_PRELOAD helper for the GPU box (no gdb there): prints a native backtrace on SIGSEGV / SIGBUS / SIGABRT.
 * gcc -shared -fPIC -O1 -o /tmp/segv_bt.so tools/segv_bt.c ; LD_PRELOAD=/tmp/segv_bt.so python ... */
#define _GNU_SOURCE
#include <execinfo.h>
#include <signal.h>
#include <string.h>
#include <unistd.h>

static void handler(int sig) {
    void* frames[64];
    const char msg[] = "\n==== native backtrace ====\n";
    write(2, msg, sizeof(msg) - 1);
    int n = backtrace(frames, 64);
    backtrace_symbols_fd(frames, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}

__attribute__((constructor)) static void install(void) {
    struct sigaction sa;
    memset(&sa, 0, sizeof(sa));
    sa.sa_handler = handler;
    sigaction(SIGSEGV, &sa, 0);
    sigaction(SIGBUS, &sa, 0);
}
