// CPU-only test harness of gamg_host.h (built with g++, no CUDA): tests/test_gamg_host_logic.py compares the
// functions gamg.cu runs on the host with the reference-pinned oracle, and the flat coarse addressing with the
// straightforward version below.  Not part of libldu_b200.so.
#include <cstring>
#include <map>

#include "gamg_host.h"

namespace {
// host: coarse owner/neighbour and the fine-face -> coarse-face map
static void coarse_addressing_simple(int nCoarse, const std::vector<int>& lower, const std::vector<int>& upper,
                              const std::vector<int>& cmap, std::vector<int>& faceMap,
                              std::vector<int>& cOwner, std::vector<int>& cNeighbour)
{
    const int nFaces = (int)lower.size();
    faceMap.assign(nFaces, 0);
    // per coarse owner: (neighbour, provisional face id) in discovery order
    std::vector<std::vector<std::pair<int, int>>> found(nCoarse);
    int nCoarseFaces = 0;
    for (int f = 0; f < nFaces; f++) {
        const int a = cmap[upper[f]], b = cmap[lower[f]];
        if (a == b) {
            faceMap[f] = -(a + 1);  // interior to a coarse cell
            continue;
        }
        const int own = std::min(a, b), nei = std::max(a, b);
        int id = -1;
        for (const auto& e : found[own])
            if (e.first == nei) {
                id = e.second;
                break;
            }
        if (id < 0) {
            id = nCoarseFaces++;
            found[own].push_back(std::make_pair(nei, id));
        }
        faceMap[f] = id;
    }
    // renumber owner-major, discovery order within an owner (GAMGAgglomerateLduAddressing.C:158-185)
    cOwner.resize(nCoarseFaces);
    cNeighbour.resize(nCoarseFaces);
    std::vector<int> renum(nCoarseFaces);
    int cf = 0;
    for (int c = 0; c < nCoarse; c++)
        for (const auto& e : found[c]) {
            cOwner[cf] = c;
            cNeighbour[cf] = e.first;
            renum[e.second] = cf++;
        }
    for (int f = 0; f < nFaces; f++)
        if (faceMap[f] >= 0) faceMap[f] = renum[faceMap[f]];
}

}  // namespace

extern "C" {

int ldu_hosttest_pair_cluster(int nFine, int nFaces, const int* lower, const int* upper, const double* w, int* cmapOut)
{
    int nCoarse = 0;
    const std::vector<int> cmap = ldu::pair_cluster(nFine, std::vector<int>(lower, lower + nFaces),
                                                    std::vector<int>(upper, upper + nFaces),
                                                    std::vector<double>(w, w + nFaces), nCoarse);
    std::memcpy(cmapOut, cmap.data(), cmap.size() * sizeof(int));
    return nCoarse;
}

// which = 0: the version gamg.cu uses, 1: the straightforward one.  Outputs sized nFaces.  Returns the number of coarse faces.
int ldu_hosttest_coarse_addressing(int which, int nCoarse, int nFaces, const int* lower, const int* upper,
                                   const int* cmap, int nFine, int* faceMapOut, int* cOwnerOut, int* cNeighbourOut)
{
    std::vector<int> faceMap, cOwner, cNeighbour;
    const std::vector<int> l(lower, lower + nFaces), u(upper, upper + nFaces), cm(cmap, cmap + nFine);
    if (which == 0) ldu::coarse_addressing(nCoarse, l, u, cm, faceMap, cOwner, cNeighbour);
    else coarse_addressing_simple(nCoarse, l, u, cm, faceMap, cOwner, cNeighbour);
    std::memcpy(faceMapOut, faceMap.data(), faceMap.size() * sizeof(int));
    std::memcpy(cOwnerOut, cOwner.data(), cOwner.size() * sizeof(int));
    std::memcpy(cNeighbourOut, cNeighbour.data(), cNeighbour.size() * sizeof(int));
    return (int)cOwner.size();
}

// faceCellsOut / faceRestrictOut sized n.  Returns the number of coarse interface faces.
int ldu_hosttest_agglomerate_interface(int myRank, int nbrRank, int n, const int* local, const int* nbr,
                                       int* faceCellsOut, int* faceRestrictOut)
{
    std::vector<int> faceCells, faceRestrict;
    ldu::agglomerate_interface(myRank, nbrRank, std::vector<int>(local, local + n), std::vector<int>(nbr, nbr + n),
                               faceCells, faceRestrict);
    std::memcpy(faceCellsOut, faceCells.data(), faceCells.size() * sizeof(int));
    std::memcpy(faceRestrictOut, faceRestrict.data(), faceRestrict.size() * sizeof(int));
    return (int)faceCells.size();
}
}
