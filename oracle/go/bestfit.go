// Package bestfit is the best-fit specification of DESIGN.md §2 written in Go, the language of
// elastic-ai/elastic-gpu-agent, so that a maintainer can produce the Go CPU number on their own
// machine (go test -bench . ./oracle/go).
//
// NOT COMPILED OR RUN IN THIS REPOSITORY'S ENVIRONMENT: there is no Go toolchain here or on the
// GPU box (SURVEY.md Appendix B), so nothing in tests/, bench.py or the reported numbers uses
// this file.  The checked oracles are oracle/bestfit_oracle.c and oracle/bestfit_np.py; this is
// the same scalar loop as the C one (SURVEY.md §8(d) asks for it to be shipped, marked so).
//
// TEST INFRASTRUCTURE ONLY, PARITY UNPINNED: the reference has no best-fit loop to restate
// (GetPreferredAllocation is a stub, pkg/plugins/base.go:94-96).  Units follow the reference:
// core in percent, 100 per card (pkg/common/const.go:4); memory in MiB
// (pkg/plugins/gpushare.go:159-168); device index = N of /dev/nvidiaN
// (pkg/operator/gpushare.go:10,32).
package bestfit

// Pick returns the feasible device with the smallest (leftover core, leftover mem, index), or -1.
func Pick(freeCore, freeMem []int32, core, mem int32) int32 {
	if core < 0 || mem < 0 {
		return -1
	}
	best := int32(-1)
	var bestLc, bestLm int32
	for d := range freeCore {
		if freeCore[d] < core || freeMem[d] < mem {
			continue
		}
		lc, lm := freeCore[d]-core, freeMem[d]-mem
		// strict "<" in index order keeps the lowest index on ties
		if best < 0 || lc < bestLc || (lc == bestLc && lm < bestLm) {
			best, bestLc, bestLm = int32(d), lc, lm
		}
	}
	return best
}

// Snapshot scores every request against the same table (DESIGN.md §2.4): indices and the demand
// each device would receive.
func Snapshot(freeCore, freeMem, reqCore, reqMem []int32, idx []int32) (deltaCore, deltaMem []int64) {
	deltaCore = make([]int64, len(freeCore))
	deltaMem = make([]int64, len(freeCore))
	for r := range reqCore {
		d := Pick(freeCore, freeMem, reqCore[r], reqMem[r])
		idx[r] = d
		if d >= 0 {
			deltaCore[d] += int64(reqCore[r])
			deltaMem[d] += int64(reqMem[r])
		}
	}
	return deltaCore, deltaMem
}

// Sequential applies allocations in order, each against the table the previous ones left
// (DESIGN.md §2.6, ALLOC events only); freeCore and freeMem are updated in place.
func Sequential(freeCore, freeMem, reqCore, reqMem []int32, idx []int32) {
	for r := range reqCore {
		d := Pick(freeCore, freeMem, reqCore[r], reqMem[r])
		idx[r] = d
		if d >= 0 {
			freeCore[d] -= reqCore[r]
			freeMem[d] -= reqMem[r]
		}
	}
}
