// oracle/spec/ngram_spec.hpp — CPU RESTATEMENT OF THE REFERENCE ALGORITHM. TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
// The product (stringsearchlib_b200/csrc) never includes, links or calls it.
//
// What it is: a clean-room restatement of serena-yu17/StringSearchLib's index()/search() path with
// DETERMINISTIC ids, templated on the character type and the gram size. It is the bit-exact target of
// the CUDA engine (gram table, postings, edges, counts, scores, ranked lists) and the only oracle for the
// surfaces the as-shipped reference does not implement (gSize != 3, wide chars) — those are
// "parity unpinned" (SURVEY.md §0.2): generalised from the literals the reference still carries.
// For narrow chars and gSize == 3 it is pinned against the compiled reference (oracle/_ref) by
// tests/test_oracle_vs_reference.py and against the golden vectors of SearchTest/test.cpp:7-18.
//
// Reference citations are relative to /root/reference/nGramSearch/.
//
// Deterministic id scheme (the reference's ids are unordered_set iteration order, nGramSearch.hpp:65):
//   ext key id : order of first occurrence of the trimmed raw key string over valid rows
//   kid        : rank of a key by (length asc, ext key id asc)  -> the tie-break of ScoreComparer
//                (nGramSearch.h:262-269) extended by "key id asc" becomes simply "kid asc"
//   term id    : rank of a normalised term by (min kid over its edges, word index of first edge)
//   gram order : lexicographic by code units
// Decisions where the reference is iteration-order dependent (documented in DESIGN.md):
//   * per-key merge is a true max over edges of (promoted ? 100 : w*s), floored at +0.0f
//     (nGramSearch.hpp:326-337 folds in hash order; differs only if some w*s > 100 or for -0.0f)
//   * wildcard score of a key = max weight over its edges (nGramSearch.hpp:358-368: last iterated wins)
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

namespace spec {

template <class C> struct CharOps;

// narrow: std::isspace/toupper in the "C" locale on unsigned char (nGramSearch.h:30-34,51-55,72-76)
template <> struct CharOps<char> {
	using U = unsigned char;
	static bool space(char c) { U u = (U)c; return u == ' ' || (u >= 9 && u <= 13); }
	static char upper(char c) { return (c >= 'a' && c <= 'z') ? char(c - 32) : c; }
	static uint32_t unit(char c) { return (U)c; }
	static constexpr int unit_bits = 8;
};
// wide: iswspace/towupper in the "C" locale are ASCII-only (SURVEY.md §8a A1 [probe]); wchar_t is 32-bit on Linux
template <> struct CharOps<wchar_t> {
	static bool space(wchar_t c) { return c == L' ' || (c >= 9 && c <= 13); }
	static wchar_t upper(wchar_t c) { return (c >= L'a' && c <= L'z') ? wchar_t(c - 32) : c; }
	static uint32_t unit(wchar_t c) { return (uint32_t)c; }
	static constexpr int unit_bits = 32;
};

struct Hit { uint32_t key; float score; };   // key = ext key id

template <class C>
class Index {
public:
	using Str = std::basic_string<C>;
	using Ops = CharOps<C>;

	uint16_t g = 3;
	bool indexed = false;               // nGramSearch.h:301 (left unset on the early return; treated as false)
	uint8_t valid[256];                 // validChar, nGramSearch.h:307-313

	// keys
	std::vector<Str> keyStr;            // by ext key id: trim(raw) — case/punctuation preserved (hpp:131-132)
	std::vector<uint32_t> kidOfExt, extOfKid;
	std::vector<uint64_t> keyFirstRow;  // by ext id: first row index (global row id for sharding)
	// terms
	std::vector<Str> termStr;           // by term id
	std::vector<uint32_t> termEdgeOff;  // CSR T+1
	std::vector<uint32_t> edgeKid;      // ascending kid inside a term
	std::vector<float> edgeW;
	std::vector<uint32_t> longTerms, shortTerms;   // ascending term id (hpp:82-85)
	// grams
	std::vector<Str> gramStr;           // sorted lexicographically
	std::vector<uint64_t> gramOff;      // CSR G+1
	std::vector<uint32_t> postings;     // term ids, ascending, unique per gram (hpp:19 inserts into a set)
	size_t longest = 0;                 // hpp:72-73

	Index() { set_default_valid(); }

	void set_default_valid()
	{
		memset(valid, 0, sizeof valid);
		const char* p = ".%$ @0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ";
		for (; *p; ++p) valid[(unsigned char)*p] = 1;
	}

	// escapeBlank (nGramSearch.h:93-98). Narrow: every byte not in validChar (incl. >= 0x80) -> ' '.
	// Wide (unpinned, nGramSearch.h:306 "in the ASCII range"): only code units < 128 are filtered.
	bool is_valid(C c) const
	{
		uint32_t u = Ops::unit(c);
		if (sizeof(C) == 1) return valid[u & 255] != 0;
		return u >= 128 ? true : valid[u] != 0;
	}
	void escape(Str& s) const { for (auto& c : s) if (!is_valid(c)) c = (C)' '; }
	static void trim(Str& s)
	{
		size_t b = 0, e = s.size();
		while (b < e && Ops::space(s[b])) ++b;
		while (e > b && Ops::space(s[e - 1])) --e;
		s = s.substr(b, e - b);
	}
	static void upper(Str& s) { for (auto& c : s) c = Ops::upper(c); }
	Str normalise(Str s) const { escape(s); trim(s); upper(s); return s; }   // hpp:136-139,153-156,372-376

	// ---------------------------------------------------------------- build (hpp:120-172, 54-108, 13-21)
	// rowBase: global index of the first row (multi-GPU key-range shards keep global first-occurrence order)
	void build(const C* const* words, uint64_t size, uint16_t rowSize, const float* weight, uint16_t gSize,
	           uint64_t rowBase = 0)
	{
		g = gSize;
		if (size < 2 || !words || rowSize == 0 || g == 0) return;                       // hpp:122
		struct RawEdge { uint32_t termTmp; uint32_t ext; float w; uint64_t word; };
		std::unordered_map<Str, uint32_t> keyIds, termIds;
		std::vector<Str> termTmpStr;
		std::vector<uint64_t> termFirstWord;
		std::vector<RawEdge> raw;
		for (uint64_t i = 0; i < size; i += rowSize) {                                  // hpp:126
			if (!words[i]) continue;                                                    // hpp:129
			Str key(words[i]);
			trim(key);
			if (key.empty()) continue;                                                  // hpp:134
			auto kit = keyIds.find(key);
			uint32_t ext;
			if (kit == keyIds.end()) {
				ext = (uint32_t)keyStr.size();
				keyIds.emplace(key, ext);
				keyStr.push_back(key);
				keyFirstRow.push_back(rowBase + i / rowSize);
			} else ext = kit->second;
			auto add_edge = [&](const Str& term, float w, uint64_t word) {
				if (w == 0.0f) return;                                                  // hpp:144,162
				auto tit = termIds.find(term);
				uint32_t t;
				if (tit == termIds.end()) {
					t = (uint32_t)termTmpStr.size();
					termIds.emplace(term, t);
					termTmpStr.push_back(term);
					termFirstWord.push_back(word);
				} else t = tit->second;
				raw.push_back({t, ext, w, word});
			};
			add_edge(normalise(key), weight ? weight[i] : 1.0f, i);                     // hpp:136-148 (may be "")
			uint64_t jend = std::min<uint64_t>(i + rowSize, size);
			for (uint64_t j = i + 1; j < jend; ++j) {                                   // hpp:150-168
				if (!words[j]) continue;
				Str term = normalise(Str(words[j]));
				if (term.empty()) continue;                                             // hpp:157
				add_edge(term, weight ? weight[j] : 1.0f, j);
			}
		}
		// kid = rank by (length, ext id)
		const uint32_t K = (uint32_t)keyStr.size();
		extOfKid.resize(K);
		for (uint32_t i = 0; i < K; ++i) extOfKid[i] = i;
		std::stable_sort(extOfKid.begin(), extOfKid.end(),
		                 [&](uint32_t a, uint32_t b) { return keyStr[a].size() < keyStr[b].size(); });
		kidOfExt.resize(K);
		for (uint32_t k = 0; k < K; ++k) kidOfExt[extOfKid[k]] = k;
		// term id = rank by (min kid over edges, first word)
		const uint32_t T = (uint32_t)termTmpStr.size();
		std::vector<uint32_t> minKid(T, 0xFFFFFFFFu);
		for (auto& e : raw) minKid[e.termTmp] = std::min(minKid[e.termTmp], kidOfExt[e.ext]);
		std::vector<uint32_t> order(T);
		for (uint32_t i = 0; i < T; ++i) order[i] = i;
		std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
			if (minKid[a] != minKid[b]) return minKid[a] < minKid[b];
			return termFirstWord[a] < termFirstWord[b];
		});
		std::vector<uint32_t> newId(T);
		termStr.resize(T);
		for (uint32_t r = 0; r < T; ++r) { newId[order[r]] = r; termStr[r] = termTmpStr[order[r]]; }
		// edges: (term, kid) unique, the LAST occurrence's weight wins (hpp:147,165)
		struct E { uint32_t t, kid; float w; uint64_t word; };
		std::vector<E> es;
		es.reserve(raw.size());
		for (auto& e : raw) es.push_back({newId[e.termTmp], kidOfExt[e.ext], e.w, e.word});
		std::stable_sort(es.begin(), es.end(), [](const E& a, const E& b) {
			if (a.t != b.t) return a.t < b.t;
			return a.kid < b.kid;
		});
		termEdgeOff.assign(T + 1, 0);
		for (size_t i = 0; i < es.size(); ++i) {
			bool last = (i + 1 == es.size()) || es[i + 1].t != es[i].t || es[i + 1].kid != es[i].kid;
			if (!last) continue;
			edgeKid.push_back(es[i].kid);
			edgeW.push_back(es[i].w);
			termEdgeOff[es[i].t + 1]++;
		}
		for (uint32_t t = 0; t < T; ++t) termEdgeOff[t + 1] += termEdgeOff[t];
		// long / short split at 2*gSize (hpp:82; nGramSearch.h:281,286), longest (hpp:72-73)
		for (uint32_t t = 0; t < T; ++t) {
			(termStr[t].size() >= (size_t)2 * g ? longTerms : shortTerms).push_back(t);
			longest = std::max(longest, termStr[t].size());
		}
		for (auto& k : keyStr) longest = std::max(longest, k.size());
		// grams of long terms: windows i = 0 .. len-g (hpp:16), posting SETS (hpp:19)
		std::vector<std::pair<Str, uint32_t>> gp;
		for (uint32_t t : longTerms) {
			const Str& s = termStr[t];
			for (size_t i = 0; i + g <= s.size(); ++i) gp.emplace_back(s.substr(i, g), t);
		}
		std::sort(gp.begin(), gp.end(), [](const auto& a, const auto& b) {
			int c = cmp_units(a.first, b.first);
			if (c) return c < 0;
			return a.second < b.second;
		});
		gp.erase(std::unique(gp.begin(), gp.end()), gp.end());
		gramOff.push_back(0);
		for (size_t i = 0; i < gp.size(); ++i) {
			if (i == 0 || gp[i].first != gp[i - 1].first) {
				if (i) gramOff.push_back(postings.size());
				gramStr.push_back(gp[i].first);
			}
			postings.push_back(gp[i].second);
		}
		if (!gp.empty()) gramOff.push_back(postings.size());
		indexed = true;                                                                 // hpp:45
	}

	static int cmp_units(const Str& a, const Str& b)
	{
		size_t n = std::min(a.size(), b.size());
		for (size_t i = 0; i < n; ++i) {
			uint32_t x = Ops::unit(a[i]), y = Ops::unit(b[i]);
			if (x != y) return x < y ? -1 : 1;
		}
		return a.size() < b.size() ? -1 : (a.size() > b.size() ? 1 : 0);
	}

	uint64_t size() const { return termStr.size(); }       // hpp:488-491 (wordMap.size())
	uint64_t lib_size() const { return gramStr.size(); }   // hpp:496-499 (ngrams.size())

	// ---------------------------------------------------------------- search (hpp:182-404)
	// stringMatch (hpp:182-222): best approximate-substring match length of q inside s.
	static size_t string_match(const Str& q, const Str& s, std::vector<uint32_t>& prev, std::vector<uint32_t>& cur)
	{
		if (q.size() == 1) {                                                            // hpp:185-191
			for (auto c : s) if (c == q[0]) return 1;
			return 0;
		}
		const size_t nq = q.size(), ns = s.size();
		prev.assign(ns + 1, 0);                                                         // free start (hpp:197)
		cur.assign(ns + 1, 0);
		for (size_t a = 0; a < nq; ++a) {
			cur[0] = (uint32_t)(a + 1);                                                 // hpp:200
			for (size_t j = 0; j < ns; ++j) {
				uint32_t cost = q[a] != s[j];
				cur[j + 1] = std::min(std::min(prev[j + 1] + 1, cur[j] + 1), prev[j] + cost);   // hpp:206-211
			}
			std::swap(prev, cur);
		}
		uint32_t mis = std::numeric_limits<uint32_t>::max();
		for (size_t j = 0; j <= ns; ++j) mis = std::min(mis, prev[j]);                  // free end (hpp:217-220)
		return nq - mis;
	}

	int64_t find_gram(const Str& gs) const
	{
		size_t lo = 0, hi = gramStr.size();
		while (lo < hi) {
			size_t mid = (lo + hi) / 2;
			if (cmp_units(gramStr[mid], gs) < 0) lo = mid + 1; else hi = mid;
		}
		return (lo < gramStr.size() && gramStr[lo] == gs) ? (int64_t)lo : -1;
	}

	// match counts of the long path for one already-normalised query (hpp:278-298); for table-level parity
	void long_counts(const Str& q, std::vector<std::pair<uint32_t, uint32_t>>& out) const
	{
		out.clear();
		if (q.size() < g) return;                                                       // hpp:281
		std::unordered_map<uint32_t, uint32_t> cnt;
		for (size_t i = 0; i + g <= q.size(); ++i) {                                    // duplicates KEPT (hpp:29-36)
			int64_t gi = find_gram(q.substr(i, g));
			if (gi < 0) continue;
			for (uint64_t p = gramOff[gi]; p < gramOff[gi + 1]; ++p) cnt[postings[p]]++;
		}
		out.assign(cnt.begin(), cnt.end());
		std::sort(out.begin(), out.end());
	}

	std::vector<Hit> search(const C* query, float threshold, uint32_t limit) const
	{
		std::vector<Hit> res;
		if (!indexed || !query) return res;                                             // hpp:417,451
		if (limit == 0) limit = (uint32_t)std::numeric_limits<int32_t>::max();          // hpp:420-421
		Str q(query);
		const uint32_t K = (uint32_t)keyStr.size();
		std::vector<float> entry(K, 0.0f);
		std::vector<uint8_t> present(K, 0);
		std::vector<uint32_t> touched;
		auto put = [&](uint32_t kid, float v) {
			if (!present[kid]) { present[kid] = 1; touched.push_back(kid); entry[kid] = v; }
			else if (v > entry[kid]) entry[kid] = v;
		};
		if (q.size() == 0 || (q.size() == 1 && q[0] == (C)'*')) {                       // wildcard, hpp:356-369
			for (uint32_t t = 0; t < termStr.size(); ++t)
				for (uint32_t e = termEdgeOff[t]; e < termEdgeOff[t + 1]; ++e) put(edgeKid[e], edgeW[e]);
		} else {
			escape(q); trim(q);
			if (q.empty()) return res;                                                  // hpp:374-375
			upper(q);
			const size_t nq = q.size();
			std::vector<std::pair<uint32_t, float>> scores;   // (term, s_t): short map first, then long map (hpp:393-394)
			if (nq < (size_t)3 * g) {                                                   // hpp:381 (literal 9 = 3*gSize)
				std::vector<uint32_t> r1, r2;
				for (uint32_t t : shortTerms)                                           // hpp:240-245
					scores.emplace_back(t, (float)string_match(q, termStr[t], r1, r2) / (float)nq);
				if (nq <= g)                                                            // hpp:247 (literal 3 = gSize)
					for (uint32_t t : longTerms)
						scores.emplace_back(t, (float)string_match(q, termStr[t], r1, r2) / (float)nq);
			}
			if (nq >= g) {                                                              // hpp:281
				std::vector<std::pair<uint32_t, uint32_t>> cnt;
				long_counts(q, cnt);
				const size_t nGrams = nq - g + 1;
				for (auto& c : cnt) scores.emplace_back(c.first, (float)c.second / (float)nGrams);   // hpp:300
			}
			for (auto& sp : scores) {                                                   // calcScore, hpp:310-341
				if (sp.second < threshold) continue;                                    // unweighted compare, hpp:315
				const bool perfect = (double)sp.second > 0.999;                         // hpp:328 (double compare)
				for (uint32_t e = termEdgeOff[sp.first]; e < termEdgeOff[sp.first + 1]; ++e) {
					float v = std::max(edgeW[e] * sp.second, 0.0f);                     // entryScore default 0 (hpp:326)
					if (v == 0.0f) v = 0.0f;                                            // canonicalise -0.0f
					if (perfect) {
						Str lib(keyStr[extOfKid[edgeKid[e]]]);                           // key NOT upper-cased, hpp:330-335
						escape(lib); trim(lib);
						if (lib == q) v = 100.0f;
					}
					put(edgeKid[e], v);
				}
			}
		}
		// order: score desc, key length asc (nGramSearch.h:262-269), then ext id asc == kid asc
		std::sort(touched.begin(), touched.end(), [&](uint32_t a, uint32_t b) {
			if (entry[a] > entry[b]) return true;
			if (entry[a] < entry[b]) return false;
			return a < b;
		});
		size_t n = std::min<size_t>(touched.size(), limit);                             // hpp:459
		res.reserve(n);
		for (size_t i = 0; i < n; ++i) res.push_back({extOfKid[touched[i]], entry[touched[i]]});
		return res;
	}
};

}   // namespace spec
