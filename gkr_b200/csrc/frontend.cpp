// Front end of the reference in native code: iden3 .r1cs / .wtns parsing and the R1CS -> expression trees ->
// layered add/mult circuit compiler, restated from
//   rust/src/convert.rs:9-10     DEPTH_LIMIT, WIDTH_LIMIT
//   rust/src/convert.rs:108-152  merge_nodes, get_k
//   rust/src/convert.rs:154-358  compile
//   rust/src/convert.rs:360-632  convert_constraints_to_nodes (the symbol-table substitution is disabled there, :565)
//   rust/src/convert.rs:793-811  input layer values from the witness
//   rust/src/aggregator.rs:399-404  how the files are read (crates r1cs-file / wtns-file = the iden3 formats)
// Output = the dense boundary of this library (gkr_layer_desc lists + input-layer values), ready for
// gkr_circuit_create / gkr_witness_eval.  Host-only C++; gkr_b200/frontend.py is the same algorithm in Python and the
// two are tested equal (tests/test_frontend.py).  Structurally equal trees are interned, so the reference's
// `next_nodes.contains(x)` / `.position(..)` (structural PartialEq, convert.rs:33-56) is an integer comparison.
#include <algorithm>
#include <array>
#include <cstdarg>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <new>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/gkr_b200.h"

namespace gkr {
void set_last_error(const char *fmt, ...);
}

namespace {

using U256 = std::array<uint64_t, 4>;
constexpr U256 kP = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
constexpr size_t kWidthLimit = 20;      // convert.rs:10

struct FrontendError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

bool geq(const U256 &a, const U256 &b) {
    for (int i = 3; i >= 0; --i)
        if (a[i] != b[i]) return a[i] > b[i];
    return true;
}
bool is_zero(const U256 &a) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
U256 neg_mod_p(const U256 &a) {          // (p - a) mod p
    if (is_zero(a)) return a;
    U256 r;
    unsigned __int128 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned __int128 d = (unsigned __int128)kP[i] - a[i] - (uint64_t)borrow;
        r[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
    return r;
}
const U256 kOne = {1, 0, 0, 0};
const U256 kMinusOne = neg_mod_p(kOne);

// ---- readers ----------------------------------------------------------------------------------------
struct Reader {
    const uint8_t *p;
    size_t n, off = 0;
    void need(uint64_t k) const {
        // off <= n is an invariant; written so that a 64-bit size taken from the file cannot wrap the comparison
        if (off > n || k > (uint64_t)(n - off)) throw FrontendError("truncated file");
    }
    uint32_t u32() {
        need(4);
        uint32_t v;
        std::memcpy(&v, p + off, 4);
        off += 4;
        return v;
    }
    uint64_t u64() {
        need(8);
        uint64_t v;
        std::memcpy(&v, p + off, 8);
        off += 8;
        return v;
    }
    U256 fe() {
        need(32);
        U256 v;
        std::memcpy(v.data(), p + off, 32);
        off += 32;
        return v;
    }
};
using Sections = std::map<uint32_t, std::pair<const uint8_t *, size_t>>;
Sections read_sections(const uint8_t *data, size_t len, const char *magic, uint32_t want_version) {
    if (len < 12 || std::memcmp(data, magic, 4) != 0) throw FrontendError(std::string("not a ") + magic + " file");
    Reader r{data, len, 4};
    const uint32_t version = r.u32(), n_sections = r.u32();
    if (version != want_version) throw FrontendError(std::string("unsupported ") + magic + " version");
    Sections out;
    for (uint32_t i = 0; i < n_sections; ++i) {
        const uint32_t ty = r.u32();
        const uint64_t size = r.u64();
        r.need(size);                                                      // rejects sizes beyond the end of the file
        out.emplace(ty, std::make_pair(data + r.off, (size_t)size));       // first section of a type wins
        r.off += (size_t)size;
    }
    return out;
}
using LinComb = std::vector<std::pair<U256, uint32_t>>;        // (coeff, wire), the tuple order of the crate
struct R1cs {
    uint32_t n_wires = 0, n_pub_out = 0, n_pub_in = 0, n_prv_in = 0;
    std::vector<std::array<LinComb, 3>> constraints;
};
R1cs parse_r1cs(const uint8_t *data, size_t len) {
    const Sections secs = read_sections(data, len, "r1cs", 1);
    if (!secs.count(1) || !secs.count(2)) throw FrontendError("r1cs file lacks the header or the constraint section");
    Reader h{secs.at(1).first, secs.at(1).second};
    if (h.u32() != 32) throw FrontendError("field size != 32: the reference reads R1csFile::<32> only");
    if (h.fe() != kP) throw FrontendError("r1cs prime is not the BN254 scalar field");
    R1cs r;
    r.n_wires = h.u32();
    r.n_pub_out = h.u32();
    r.n_pub_in = h.u32();
    r.n_prv_in = h.u32();
    h.u64();
    const uint32_t n_constraints = h.u32();
    Reader b{secs.at(2).first, secs.at(2).second};
    // counts come from the file: bound them by what the section can hold before sizing anything from them
    // (a constraint is at least three empty term lists = 12 bytes, a term 36 bytes)
    if ((uint64_t)n_constraints * 12 > b.n) throw FrontendError("constraint count exceeds the constraint section");
    r.constraints.resize(n_constraints);
    for (auto &c : r.constraints)
        for (auto &lc : c) {
            const uint32_t n = b.u32();
            if ((uint64_t)n * 36 > b.n - b.off) throw FrontendError("term count exceeds the constraint section");
            lc.reserve(n);
            for (uint32_t i = 0; i < n; ++i) {
                const uint32_t wire = b.u32();
                const U256 coeff = b.fe();
                if (geq(coeff, kP) || wire >= r.n_wires) throw FrontendError("constraint term out of range");
                lc.emplace_back(coeff, wire);
            }
        }
    return r;
}
std::vector<U256> parse_wtns(const uint8_t *data, size_t len) {
    const Sections secs = read_sections(data, len, "wtns", 2);
    if (!secs.count(1) || !secs.count(2)) throw FrontendError("wtns file lacks the header or the data section");
    Reader h{secs.at(1).first, secs.at(1).second};
    if (h.u32() != 32) throw FrontendError("field size != 32: the reference reads WtnsFile::<32> only");
    if (h.fe() != kP) throw FrontendError("wtns prime is not the BN254 scalar field");
    const uint32_t n = h.u32();
    Reader d{secs.at(2).first, secs.at(2).second};
    if ((uint64_t)n * 32 > d.n) throw FrontendError("witness count exceeds the data section");
    std::vector<U256> w(n);
    for (auto &v : w) {
        v = d.fe();
        if (geq(v, kP)) throw FrontendError("witness value out of range");     // from_repr(..).unwrap(), convert.rs:806
    }
    return w;
}

// ---- interned expression trees ------------------------------------------------------------------------
enum Kind : uint8_t { kValue = 0, kVar = 1, kAdd = 2, kMult = 3 };
struct Node {
    Kind kind;
    U256 value;          // kValue
    uint32_t var;        // kVar
    int left, right;     // kAdd / kMult
    int depth;           // a leaf has depth 1 (convert.rs:85-89)
};
struct KeyHash {
    size_t operator()(const std::array<uint64_t, 7> &k) const {
        uint64_t h = 0x9e3779b97f4a7c15ULL;
        for (uint64_t v : k) h = (h ^ v) * 0xff51afd7ed558ccdULL + (h >> 29);
        return (size_t)h;
    }
};
struct Pool {
    std::vector<Node> nodes;
    std::unordered_map<std::array<uint64_t, 7>, int, KeyHash> index;
    int intern(const Node &n) {
        const std::array<uint64_t, 7> key = {(uint64_t)n.kind, n.value[0], n.value[1], n.value[2], n.value[3],
                                            (uint64_t)n.var, ((uint64_t)(uint32_t)n.left << 32) | (uint32_t)n.right};
        auto it = index.find(key);
        if (it != index.end()) return it->second;
        nodes.push_back(n);
        index.emplace(key, (int)nodes.size() - 1);
        return (int)nodes.size() - 1;
    }
    int value(const U256 &v) { return intern(Node{kValue, v, 0, -1, -1, 1}); }
    int var(uint32_t x) { return intern(Node{kVar, U256{0, 0, 0, 0}, x, -1, -1, 1}); }
    int op(Kind k, int l, int r) {
        return intern(Node{k, U256{0, 0, 0, 0}, 0, l, r, std::max(nodes[l].depth, nodes[r].depth) + 1});
    }
    // coeff * x_wire, the multiplication elided when coeff == 1
    int term(const U256 &coeff, uint32_t wire) { return coeff == kOne ? var(wire) : op(kMult, value(coeff), var(wire)); }
};

// convert.rs:108-139; on an empty list the reference recurses forever: rejected
int merge_nodes(Pool &pool, const std::vector<int> &nodes) {
    if (nodes.empty())
        throw FrontendError("empty linear combination: the reference does not terminate on it (merge_nodes, convert.rs:108-139)");
    if (nodes.size() == 1) return nodes[0];
    std::vector<int> next;
    for (size_t i = 0; i + 1 < nodes.size(); i += 2) next.push_back(pool.op(kAdd, nodes[i], nodes[i + 1]));
    if (nodes.size() % 2 == 1) return pool.op(kAdd, merge_nodes(pool, next), nodes.back());
    return merge_nodes(pool, next);
}
uint32_t get_k(size_t n) {          // convert.rs:141-152
    if (n == 0) throw FrontendError("get_k(0)");
    uint32_t k = 0;
    while (((size_t)1 << k) < n) ++k;
    return k;
}
std::pair<int, int> count_mult(const LinComb &lc) {       // convert.rs:363-378
    int a = 0, b = 0;
    for (const auto &t : lc) {
        if (t.first == kOne) ++b;
        else if (t.first == kMinusOne) ++a;
        else { ++a; ++b; }
    }
    return {a, b};
}
// convert.rs:360-632 with the (disabled) symbol table always empty: one tree per constraint,
//   neg = false:  A * B + (-C)      neg = true:  (-A) * B + C
std::vector<std::vector<int>> constraints_to_nodes(Pool &pool, const R1cs &r) {
    std::vector<std::vector<int>> groups;
    for (const auto &c : r.constraints) {
        const LinComb &a = c[0], &b = c[1], &cc = c[2];
        const auto ca = count_mult(a), cb = count_mult(b), c3 = count_mult(cc);
        const bool neg = ca.first + cb.first + c3.second > ca.second + cb.second + c3.first;
        auto negated = [&](const LinComb &lc) {          // -lc, with -(-1) * x written as x
            std::vector<int> out;
            for (const auto &t : lc)
                out.push_back(t.first == kMinusOne ? pool.var(t.second) : pool.op(kMult, pool.value(neg_mod_p(t.first)), pool.var(t.second)));
            return out;
        };
        auto plain = [&](const LinComb &lc) {
            std::vector<int> out;
            for (const auto &t : lc) out.push_back(pool.term(t.first, t.second));
            return out;
        };
        const std::vector<int> node_a = neg ? negated(a) : plain(a), node_b = plain(b);
        if (!node_a.empty() && !node_b.empty()) {
            const int ab = pool.op(kMult, merge_nodes(pool, node_a), merge_nodes(pool, node_b));
            const std::vector<int> node_c = neg ? plain(cc) : negated(cc);
            groups.push_back({pool.op(kAdd, ab, merge_nodes(pool, node_c))});
        } else {
            groups.push_back({merge_nodes(pool, {})});       // the reference merges node_c before filling it (:620-623)
        }
    }
    return groups;
}

struct LayerOut {
    std::vector<uint8_t> type;
    std::vector<uint32_t> left, right;
};
struct SubCircuit {
    std::vector<LayerOut> layers;
    std::vector<uint32_t> k;               // k_0 .. k_depth
    std::vector<gkr_fr> input_values;      // 2^k_depth canonical values
    std::vector<gkr_layer_desc> descs;
};

// convert.rs:154-358
std::vector<SubCircuit> compile(Pool &pool, std::vector<std::vector<int>> groups, const std::vector<U256> &witness) {
    auto height_of = [&](const std::vector<int> &g) {
        int h = 0;
        for (int n : g) h = std::max(h, pool.nodes[n].depth);
        return h;
    };
    std::stable_sort(groups.begin(), groups.end(),
                     [&](const std::vector<int> &x, const std::vector<int> &y) { return height_of(x) < height_of(y); });
    while (groups.size() > kWidthLimit) {
        std::vector<std::vector<int>> merged;
        for (size_t i = 0; i + 1 < groups.size(); i += 2) {
            std::vector<int> g = groups[i];
            g.insert(g.end(), groups[i + 1].begin(), groups[i + 1].end());
            merged.push_back(std::move(g));
        }
        if (groups.size() % 2 == 1) merged.push_back(groups.back());
        groups.swap(merged);
    }
    const int zero = pool.value(U256{0, 0, 0, 0});
    std::vector<SubCircuit> out;
    for (const auto &one_circuit : groups) {
        SubCircuit sc;
        const int height = height_of(one_circuit);
        if (height == 0) {               // convert.rs:193-195 returns at once with no input lists: nothing to prove
            out.clear();
            return out;
        }
        std::vector<int> current = one_circuit;
        for (int d = 0; d <= height; ++d) {
            const uint32_t k = get_k(current.size());
            current.resize((size_t)1 << k, zero);
            sc.k.push_back(k);
            if (d == height) {
                for (int n : current) {
                    const Node &nd = pool.nodes[n];
                    U256 v;
                    if (nd.kind == kValue) v = nd.value;
                    else if (nd.kind == kVar) {
                        if (nd.var >= witness.size()) throw FrontendError("a wire of the circuit is not in the witness");
                        v = witness[nd.var];
                    } else throw FrontendError("input layer holds an operation");
                    gkr_fr f;
                    std::memcpy(f.l, v.data(), 32);
                    sc.input_values.push_back(f);
                }
                break;
            }
            std::vector<int> next;
            std::unordered_map<int, uint32_t> first_pos, used;
            int64_t zero_index = -1;
            LayerOut L;
            auto push = [&](int n) {
                next.push_back(n);
                first_pos.emplace(n, (uint32_t)next.size() - 1);        // keeps the first occurrence
                return (uint32_t)next.size() - 1;
            };
            for (int n : current) {
                const Node nd = pool.nodes[n];
                if (nd.kind == kAdd || nd.kind == kMult) {
                    if (d == height - 1) throw FrontendError("Unsupported");           // convert.rs:219-221
                    L.type.push_back(nd.kind == kAdd ? 0 : 1);
                    auto itl = first_pos.find(nd.left);
                    const uint32_t li = itl != first_pos.end() ? itl->second : push(nd.left);
                    auto itr = first_pos.find(nd.right);
                    const uint32_t ri = itr != first_pos.end() ? itr->second : push(nd.right);
                    L.left.push_back(li);
                    L.right.push_back(ri);
                } else {
                    L.type.push_back(0);
                    auto it = used.find(n);
                    if (it != used.end()) {
                        L.left.push_back(it->second);
                        L.right.push_back((uint32_t)zero_index);
                        continue;
                    }
                    if (zero_index < 0) zero_index = push(zero);
                    if (n == zero) {
                        used.emplace(n, (uint32_t)zero_index);
                        L.left.push_back((uint32_t)zero_index);
                        L.right.push_back((uint32_t)zero_index);
                    } else {
                        used.emplace(n, (uint32_t)next.size());
                        L.left.push_back((uint32_t)next.size());
                        L.right.push_back((uint32_t)zero_index);
                        push(n);
                    }
                }
            }
            sc.layers.push_back(std::move(L));
            current.swap(next);
        }
        out.push_back(std::move(sc));
    }
    for (auto &sc : out) {
        sc.descs.resize(sc.layers.size());
        for (size_t i = 0; i < sc.layers.size(); ++i) {
            gkr_layer_desc &d = sc.descs[i];
            d.k_out = sc.k[i];
            d.k_in = sc.k[i + 1];
            d.n_gates = (uint32_t)sc.layers[i].type.size();
            d.type = sc.layers[i].type.data();
            d.left = sc.layers[i].left.data();
            d.right = sc.layers[i].right.data();
        }
    }
    return out;
}

}  // namespace

struct gkr_frontend {
    std::vector<SubCircuit> circuits;
    uint32_t n_pub = 0;
    // Output of the reference (convert.rs:634-667): wire i+1 -> (witness value, name from the .sym file), i < n_pub
    std::vector<std::string> out_names;
    std::vector<gkr_fr> out_values;
};

namespace {
// parse_sym (convert.rs:851-871): the name after `main.` in the 4th comma-separated column of the first num_public
// lines.  The reference indexes l[3] and name_main[1] unchecked (it panics on a malformed line); here that is an error.
std::vector<std::string> parse_sym(const char *text, size_t len, uint32_t num_public) {
    std::vector<std::string> res;
    if (num_public == 0) return res;
    size_t pos = 0;
    while (pos < len && res.size() < num_public) {
        size_t eol = pos;
        while (eol < len && text[eol] != '\n') ++eol;
        size_t end = eol;
        if (end > pos && text[end - 1] == '\r') --end;          // str::lines() strips \r\n as well
        const std::string line(text + pos, end - pos);
        pos = eol + 1;
        size_t col = 0, start = 0;
        for (int c = 0; c < 3; ++c) {
            col = line.find(',', start);
            if (col == std::string::npos) throw FrontendError("sym line has fewer than four columns");
            start = col + 1;
        }
        const size_t stop = line.find(',', start);
        const std::string field = line.substr(start, stop == std::string::npos ? std::string::npos : stop - start);
        const size_t dot = field.find('.');
        if (dot == std::string::npos) throw FrontendError("sym name has no component after `main.`");
        const size_t dot2 = field.find('.', dot + 1);
        res.push_back(field.substr(dot + 1, dot2 == std::string::npos ? std::string::npos : dot2 - dot - 1));
    }
    return res;
}
}  // namespace

static int frontend_compile(const uint8_t *r1cs, size_t r1cs_len, const uint8_t *wtns, size_t wtns_len, const char *sym,
                            size_t sym_len, gkr_frontend **out) {
    if (!r1cs || !wtns || !out) return GKR_ERR_INVALID;
    *out = nullptr;
    try {
        std::unique_ptr<gkr_frontend> fe(new gkr_frontend());
        const R1cs r = parse_r1cs(r1cs, r1cs_len);
        const std::vector<U256> w = parse_wtns(wtns, wtns_len);
        Pool pool;
        fe->circuits = compile(pool, constraints_to_nodes(pool, r), w);
        fe->n_pub = r.n_pub_in + r.n_pub_out;
        if (sym) {
            fe->out_names = parse_sym(sym, sym_len, fe->n_pub);
            // make_output (convert.rs:653-667) reads witness[i + 1] for every name
            if (fe->out_names.size() + 1 > w.size()) throw FrontendError("witness is shorter than the public signals");
            fe->out_values.resize(fe->out_names.size());
            for (size_t i = 0; i < fe->out_names.size(); ++i) std::memcpy(&fe->out_values[i], w[i + 1].data(), 32);
        }
        *out = fe.release();
        return GKR_OK;
    } catch (const FrontendError &e) {
        gkr::set_last_error("front end: %s", e.what());
        return GKR_ERR_INVALID;
    } catch (const std::bad_alloc &) {
        gkr::set_last_error("front end: out of memory");
        return GKR_ERR_OOM;
    } catch (const std::exception &e) {               // nothing may unwind through the C boundary
        gkr::set_last_error("front end: %s", e.what());
        return GKR_ERR_INVALID;
    } catch (...) {
        gkr::set_last_error("front end: unknown failure");
        return GKR_ERR_INTERNAL;
    }
}
extern "C" int gkr_frontend_compile(const uint8_t *r1cs, size_t r1cs_len, const uint8_t *wtns, size_t wtns_len,
                                    gkr_frontend **out) {
    return frontend_compile(r1cs, r1cs_len, wtns, wtns_len, nullptr, 0, out);
}
extern "C" int gkr_frontend_compile_sym(const uint8_t *r1cs, size_t r1cs_len, const uint8_t *wtns, size_t wtns_len,
                                        const char *sym, size_t sym_len, gkr_frontend **out) {
    if (!sym) return GKR_ERR_INVALID;
    return frontend_compile(r1cs, r1cs_len, wtns, wtns_len, sym, sym_len, out);
}
extern "C" uint32_t gkr_frontend_n_outputs(const gkr_frontend *fe) { return fe ? (uint32_t)fe->out_names.size() : 0; }
extern "C" int gkr_frontend_output(const gkr_frontend *fe, uint32_t i, uint32_t *wire, gkr_fr *value, const char **name) {
    if (!fe || i >= fe->out_names.size() || !wire || !value || !name) return GKR_ERR_INVALID;
    *wire = i + 1;
    *value = fe->out_values[i];
    *name = fe->out_names[i].c_str();
    return GKR_OK;
}
extern "C" uint32_t gkr_frontend_n_circuits(const gkr_frontend *fe) { return fe ? (uint32_t)fe->circuits.size() : 0; }
extern "C" uint32_t gkr_frontend_n_public(const gkr_frontend *fe) { return fe ? fe->n_pub : 0; }
extern "C" int gkr_frontend_circuit(const gkr_frontend *fe, uint32_t i, uint32_t *n_layers, const gkr_layer_desc **layers,
                                    uint32_t *input_k, const gkr_fr **input_values) {
    if (!fe || i >= fe->circuits.size() || !n_layers || !layers || !input_k || !input_values) return GKR_ERR_INVALID;
    const SubCircuit &sc = fe->circuits[i];
    if (sc.layers.empty()) {
        gkr::set_last_error("front end: sub-circuit %u is empty", i);
        return GKR_ERR_INVALID;
    }
    *n_layers = (uint32_t)sc.layers.size();
    *layers = sc.descs.data();
    *input_k = sc.k.back();
    *input_values = sc.input_values.data();
    return GKR_OK;
}
extern "C" void gkr_frontend_destroy(gkr_frontend *fe) { delete fe; }
