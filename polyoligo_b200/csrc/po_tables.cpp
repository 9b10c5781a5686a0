// po_tables.cpp -- loads thermodynamic parameters in libprimer3's primer3_config
// file format into the device table layout (PoTables).
//
// File format (upstream thal.c get_thermodynamic_values and its helpers):
//   stack.ds/.dh, stackmm.ds/.dh, tstack_tm_inf.ds + tstack.dh, tstack2.ds/.dh:
//       256 numbers (or "inf"), nested A,C,G,T in [a][b][c][d] order
//   dangle.ds/.dh: 64 numbers for dangle3 in (a, c, x) order, then 64 for dangle5
//       in (a, c, y) order
//   loops.ds/.dh: 30 lines "size interior bulge hairpin"
//   triloop.*, tetraloop.*: "SEQ value" per line
// Non-finite entries become {dH = +inf, dS = -1}, as upstream stores them.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

#include "po_common.h"

namespace {

struct Embedded {
    const char* name;
    const char* text;
};
const Embedded kEmbedded[] = {
#include "default_params.inc"
    {nullptr, nullptr}};

bool read_text(const char* dir, const std::string& name, std::string* text, std::string* err) {
    if (dir == nullptr) {
        for (const Embedded* e = kEmbedded; e->name; ++e)
            if (name == e->name) {
                *text = e->text;
                return true;
            }
        *err = "no embedded parameter file " + name;
        return false;
    }
    std::ifstream f(std::string(dir) + "/" + name, std::ios::binary);
    if (!f) {
        *err = "cannot open " + std::string(dir) + "/" + name;
        return false;
    }
    std::stringstream ss;
    ss << f.rdbuf();
    *text = ss.str();
    return true;
}

bool tokens_of(const char* dir, const std::string& name, std::vector<std::string>* tok, std::string* err) {
    std::string text;
    if (!read_text(dir, name, &text, err)) return false;
    std::istringstream is(text);
    std::string t;
    tok->clear();
    while (is >> t) tok->push_back(t);
    return true;
}

double to_double(const std::string& t) {
    if (t == "inf") return INFINITY;
    return std::strtod(t.c_str(), nullptr);
}

double2 make_entry(double h, double s) {
    double2 v;
    if (!std::isfinite(h) || !std::isfinite(s)) {
        v.x = INFINITY;
        v.y = -1.0;
    } else {
        v.x = h;
        v.y = s;
    }
    return v;
}

bool load_numeric(const char* dir, const char* fs, const char* fh, size_t n, double2* out, std::string* err) {
    std::vector<std::string> ts, th;
    if (!tokens_of(dir, fs, &ts, err) || !tokens_of(dir, fh, &th, err)) return false;
    if (ts.size() < n || th.size() < n) {
        *err = std::string("parameter file too short: ") + fs + " / " + fh;
        return false;
    }
    for (size_t k = 0; k < n; ++k) out[k] = make_entry(to_double(th[k]), to_double(ts[k]));
    return true;
}

int code_of(char ch) {
    switch (ch) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
    }
    return -1;
}

// union of the .dh and .ds lists; a key missing from one file contributes 0 there
bool load_bonus(const char* dir, const char* fs, const char* fh, int k, int32_t* keys, double2* vals,
                int32_t* n_out, std::string* err) {
    std::map<int32_t, double2> m;
    for (int pass = 0; pass < 2; ++pass) {
        std::vector<std::string> t;
        if (!tokens_of(dir, pass == 0 ? fh : fs, &t, err)) return false;
        if (t.size() % 2) {
            *err = std::string("odd token count in ") + (pass == 0 ? fh : fs);
            return false;
        }
        for (size_t i = 0; i < t.size(); i += 2) {
            if ((int)t[i].size() != k) {
                *err = std::string("bad loop key '") + t[i] + "' in " + (pass == 0 ? fh : fs);
                return false;
            }
            int32_t key = 0;
            bool acgt = true;
            for (int c = 0; c < k; ++c) {
                int code = code_of(t[i][c]);
                if (code < 0) acgt = false;
                else key |= code << (2 * c);
            }
            if (!acgt) continue;  // a loop containing N can never match a 2-bit oligo
            auto it = m.find(key);
            if (it == m.end()) {
                double2 z;
                z.x = 0.0;
                z.y = 0.0;
                it = m.insert({key, z}).first;
            }
            if (pass == 0) it->second.x = to_double(t[i + 1]);
            else it->second.y = to_double(t[i + 1]);
        }
    }
    if (m.size() > PO_MAX_BONUS) {
        *err = std::string("too many entries in ") + fh;
        return false;
    }
    int n = 0;
    for (auto& kv : m) {  // std::map iterates in ascending key order
        keys[n] = kv.first;
        vals[n] = kv.second;
        ++n;
    }
    *n_out = n;
    return true;
}

}  // namespace

bool po_load_tables(const char* dir, PoTables* T, std::string* err) {
    std::memset(T, 0, sizeof(*T));
    if (!load_numeric(dir, "stack.ds", "stack.dh", 256, T->stack, err)) return false;
    if (!load_numeric(dir, "stackmm.ds", "stackmm.dh", 256, T->stackint2, err)) return false;
    if (!load_numeric(dir, "tstack_tm_inf.ds", "tstack.dh", 256, T->tstack, err)) return false;
    if (!load_numeric(dir, "tstack2.ds", "tstack2.dh", 256, T->tstack2, err)) return false;

    // dangle files: (a, c, x) order on disk; device layout dangle3[a][x][c], dangle5[a][c][y]
    std::vector<double2> d(128);
    if (!load_numeric(dir, "dangle.ds", "dangle.dh", 128, d.data(), err)) return false;
    for (int a = 0; a < 4; ++a)
        for (int c = 0; c < 4; ++c)
            for (int x = 0; x < 4; ++x) {
                T->dangle3[(a * 4 + x) * 4 + c] = d[(a * 4 + c) * 4 + x];
                T->dangle5[(a * 4 + c) * 4 + x] = d[64 + (a * 4 + c) * 4 + x];
            }

    std::vector<std::string> ls, lh;
    if (!tokens_of(dir, "loops.ds", &ls, err) || !tokens_of(dir, "loops.dh", &lh, err)) return false;
    if (ls.size() < 120 || lh.size() < 120) {
        *err = "loops.ds / loops.dh too short";
        return false;
    }
    for (int k = 0; k < 30; ++k) {
        // upstream keeps loop entries verbatim (no inf -> -1 rewrite)
        T->interior[k].x = to_double(lh[4 * k + 1]);
        T->interior[k].y = to_double(ls[4 * k + 1]);
        T->bulge[k].x = to_double(lh[4 * k + 2]);
        T->bulge[k].y = to_double(ls[4 * k + 2]);
        T->hairpin[k].x = to_double(lh[4 * k + 3]);
        T->hairpin[k].y = to_double(ls[4 * k + 3]);
    }
    if (!load_bonus(dir, "triloop.ds", "triloop.dh", 5, T->tri_key, T->tri_val, &T->n_tri, err)) return false;
    if (!load_bonus(dir, "tetraloop.ds", "tetraloop.dh", 6, T->tetra_key, T->tetra_val, &T->n_tetra, err))
        return false;
    return true;
}
