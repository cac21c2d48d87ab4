// shim/shim_ctx.h -- what the three reference-side bindings (shim_fwd2d1.cc, shim_calcdist.cc, shim_alignc.cc,
// shim_fwd2b1.cc) share: a process-wide POOL of library contexts and the policy for calls the library refuses.
//
// Contexts.  The reference reaches the DP from short-lived pthread workers: Prrn::best_of_n creates and joins
// no_thread workers per refinement cycle (src/prrn5.cc:606-611), CalcServer::MasterWorker does the same per run
// (src/calcserv.h:436-457).  A context per calling thread would therefore be created (~1 s of CUDA set-up, streams,
// events, grow-only workspaces, pinned staging) and abandoned on every cycle.  Instead a call checks a context OUT
// of a mutex-protected free list and hands it back when it returns: at most as many contexts as there were calls
// in flight at the same time ever exist, and they live until the process ends (pg_destroy at exit).
//
// Refused calls.  The library has no CPU fallback, and neither have the shims by default: a call the CUDA path does
// not take (rectangle -A, caller-supplied window, local / Smith-Waterman modes, naive groups with nil ends or more
// than 32 members) is fatal(), the reference's own error convention.  PRRN_GPU_ALLOW_REF=1 opts in to running
// exactly those calls on the reference's own code (its unmodified function under another name), each announced on
// stderr the first time its kind occurs and counted under PRRN_GPU_STATS.
#ifndef PRRN_SHIM_CTX_H
#define PRRN_SHIM_CTX_H

#include "prrn_gpu.h"

#include <mutex>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

extern void	fatal(const char* format,...);		// src/clib.h

struct PgCtxPool {
	std::mutex	mu;
	std::vector<pg_context*>	idle;
	int	created;
	PgCtxPool() : created(0) {}
	~PgCtxPool() {for (pg_context* c: idle) pg_destroy(c);}
};

inline PgCtxPool& pg_ctx_pool()
{
	static PgCtxPool	pool;		// one per process (inline function: the same object in every shim)
	return pool;
}

// RAII lease of one context for the duration of one library call
struct PgLease {
	pg_context*	c;
	PgLease() : c(0) {
	    PgCtxPool&	pool = pg_ctx_pool();
	    {
		std::lock_guard<std::mutex>	lk(pool.mu);
		if (!pool.idle.empty()) {c = pool.idle.back(); pool.idle.pop_back();}
		else ++pool.created;
	    }
	    if (!c && pg_create(0, &c) != PG_OK) fatal("prrn_gpu: %s\n", pg_last_error(0));
	}
	~PgLease() {
	    PgCtxPool&	pool = pg_ctx_pool();
	    std::lock_guard<std::mutex>	lk(pool.mu);
	    pool.idle.push_back(c);
	}
	operator pg_context*() const {return c;}
};

inline int pg_ctx_created()
{
	PgCtxPool&	pool = pg_ctx_pool();
	std::lock_guard<std::mutex>	lk(pool.mu);
	return pool.created;
}

// A call the library does not take: fatal() unless PRRN_GPU_ALLOW_REF=1, then the caller runs the reference's code
inline void pg_refused(const char* who, const char* what)
{
	static const bool	allow = getenv("PRRN_GPU_ALLOW_REF") && getenv("PRRN_GPU_ALLOW_REF")[0] == '1';
	if (!allow)
	    fatal("prrn_gpu %s: %s is not built on the GPU path and there is no CPU fallback "
		"(PRRN_GPU_ALLOW_REF=1 runs such calls on the reference's own code)\n", who, what);
	static std::mutex	mu;
	static std::vector<const char*>	seen;
	std::lock_guard<std::mutex>	lk(mu);
	for (const char* s: seen) if (s == what) return;
	seen.push_back(what);
	fprintf(stderr, "prrn_gpu %s: %s left on the reference's own code (PRRN_GPU_ALLOW_REF=1)\n", who, what);
}

#endif
