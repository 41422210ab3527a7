// Host side of the GAMG hierarchy: one level of pairwise clustering and the coarse addressing it implies
// (pairGAMGAgglomerate.C:36-198, GAMGAgglomerateLduAddressing.C:34-214).  Pure C++ so that the same code that
// gamg.cu runs is exercised on the CPU against the reference-pinned oracle (gamg_host_test.cpp,
// tests/test_gamg_host_logic.py).
#pragma once

#include <algorithm>
#include <map>
#include <utility>
#include <vector>

namespace ldu {

constexpr double kScalarGreat = 1.0e+15;   // primitives/Scalar/doubleScalar/doubleScalar.H:54

// ---------------------------------------------------------------------------
// host: one level of pairwise clustering
// ---------------------------------------------------------------------------
inline std::vector<int> pair_cluster(int nFine, const std::vector<int>& lower, const std::vector<int>& upper,
                                     const std::vector<double>& w, int& nCoarse)
{
    const int nFaces = (int)lower.size();
    // faces around each cell: first the faces where the cell is the neighbour,
    // then those where it is the owner (the order the reference scans them in)
    std::vector<int> start(nFine + 1, 0);
    for (int f = 0; f < nFaces; f++) {
        start[upper[f] + 1]++;
        start[lower[f] + 1]++;
    }
    for (int c = 0; c < nFine; c++) start[c + 1] += start[c];
    std::vector<int> cellFaces(2 * (size_t)nFaces), fill(start.begin(), start.end() - 1);
    for (int f = 0; f < nFaces; f++) cellFaces[fill[upper[f]]++] = f;
    for (int f = 0; f < nFaces; f++) cellFaces[fill[lower[f]]++] = f;

    std::vector<int> cmap(nFine, -1);
    nCoarse = 0;
    for (int c = 0; c < nFine; c++) {
        if (cmap[c] >= 0) continue;
        int match = -1;
        double best = -kScalarGreat;
        for (int k = start[c]; k < start[c + 1]; k++) {
            const int f = cellFaces[k];
            if (cmap[upper[f]] < 0 && cmap[lower[f]] < 0 && w[f] > best) {
                match = f;
                best = w[f];
            }
        }
        if (match >= 0) {  // new pair
            cmap[upper[match]] = nCoarse;
            cmap[lower[match]] = nCoarse;
            nCoarse++;
            continue;
        }
        // no free neighbour: join the cluster across the heaviest face
        int cmatch = -1;
        best = -kScalarGreat;
        for (int k = start[c]; k < start[c + 1]; k++) {
            const int f = cellFaces[k];
            if (w[f] > best) {
                cmatch = f;
                best = w[f];
            }
        }
        if (cmatch >= 0) cmap[c] = std::max(cmap[upper[cmatch]], cmap[lower[cmatch]]);
    }
    for (int c = 0; c < nFine; c++)
        if (cmap[c] < 0) cmap[c] = nCoarse++;
    // the reference reverses the cluster numbering (pairGAMGAgglomerate.C:186-195)
    for (int c = 0; c < nFine; c++) cmap[c] = nCoarse - 1 - cmap[c];
    return cmap;
}

// host: coarse owner/neighbour and the fine-face -> coarse-face map.  Coarse faces are numbered owner-major and,
// within an owner, in the order the fine faces discover them (GAMGAgglomerateLduAddressing.C:158-185).  Flat
// arrays: the neighbours an owner has found so far live in a segment sized by the number of its fine faces (an
// upper bound), searched linearly -- a coarse cell has a handful of neighbours; no allocation per coarse cell
// (the first version kept a vector per coarse cell: a million small vectors on a 2M-cell mesh).
inline void coarse_addressing(int nCoarse, const std::vector<int>& lower, const std::vector<int>& upper,
                              const std::vector<int>& cmap, std::vector<int>& faceMap,
                              std::vector<int>& cOwner, std::vector<int>& cNeighbour)
{
    const int nFaces = (int)lower.size();
    faceMap.assign(nFaces, 0);
    std::vector<int> seg(nCoarse + 1, 0);
    for (int f = 0; f < nFaces; f++) {
        const int a = cmap[upper[f]], b = cmap[lower[f]];
        if (a != b) seg[std::min(a, b) + 1]++;
    }
    for (int c = 0; c < nCoarse; c++) seg[c + 1] += seg[c];
    std::vector<int> nbr(seg[nCoarse]), id(seg[nCoarse]), len(nCoarse, 0);
    int nCoarseFaces = 0;
    for (int f = 0; f < nFaces; f++) {
        const int a = cmap[upper[f]], b = cmap[lower[f]];
        if (a == b) {
            faceMap[f] = -(a + 1);  // interior to a coarse cell
            continue;
        }
        const int own = std::min(a, b), nei = std::max(a, b);
        const int s0 = seg[own];
        int k = 0;
        while (k < len[own] && nbr[s0 + k] != nei) k++;
        if (k == len[own]) {
            nbr[s0 + k] = nei;
            id[s0 + k] = nCoarseFaces++;
            len[own]++;
        }
        faceMap[f] = id[s0 + k];
    }
    cOwner.resize(nCoarseFaces);
    cNeighbour.resize(nCoarseFaces);
    std::vector<int> renum(nCoarseFaces);
    int cf = 0;
    for (int c = 0; c < nCoarse; c++)
        for (int k = 0; k < len[c]; k++) {
            cOwner[cf] = c;
            cNeighbour[cf] = nbr[seg[c] + k];
            renum[id[seg[c] + k]] = cf++;
        }
    for (int f = 0; f < nFaces; f++)
        if (faceMap[f] >= 0) faceMap[f] = renum[faceMap[f]];
}

// coarse processor interface of one pairing step (processorGAMGInterface.C:47-126): the faces of the interface are
// merged by the pair (coarse cell here, coarse cell there), keyed from the side of the lower rank so that both
// sides number the coarse faces alike; local / nbr = the coarse cell of each fine face on this side / the other side
inline void agglomerate_interface(int myRank, int nbrRank, const std::vector<int>& local,
                                  const std::vector<int>& nbr, std::vector<int>& faceCells,
                                  std::vector<int>& faceRestrict)
{
    std::map<std::pair<int, int>, int> seen;
    faceCells.clear();
    faceRestrict.resize(local.size());
    for (size_t i = 0; i < local.size(); i++) {
        const std::pair<int, int> key = (myRank < nbrRank) ? std::make_pair(local[i], nbr[i])
                                                           : std::make_pair(nbr[i], local[i]);
        auto it = seen.find(key);
        if (it == seen.end()) {
            const int id = (int)faceCells.size();
            seen.emplace(key, id);
            faceCells.push_back(local[i]);
            faceRestrict[i] = id;
        } else {
            faceRestrict[i] = it->second;
        }
    }
}

}  // namespace ldu
