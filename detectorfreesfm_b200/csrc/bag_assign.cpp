// bag assignment + chunking of feature tracks for the multi-view refinement matcher -- a native
// restatement of src/post_optimization/data_construct/construct_matching_data.py:10-163 (FeatureTrackStatus), :202-224 (chunk_bags),
// :226-261 (assign_bags) and src/utils/ray_utils.py:100-108 (chunks_balance).
//
// The reference is a sequential greedy loop over Python lists / sets with one np.argmax over ALL tracks per bag (quadratic: hours at 1e6
// tracks).  Its result depends on the ITERATION ORDER of CPython sets of image ids (`list(set(a) - {b})`, `list(exclude_img_ids)[:quota]`,
// `list(common_img_ids)` ...), so a drop-in must reproduce that order: PySet below is a work-alike of CPython's set for small non-negative
// integers (hash(n) == n): open addressing, LINEAR_PROBES = 9, PERTURB_SHIFT = 5, growth at fill*5 >= mask*3 to the next power of two above
// 4*used (2*used beyond 50000), the set_merge / set_difference / set_intersection strategies of Objects/setobject.c (3.7 .. 3.12).  It is
// pinned against the running interpreter (tests/test_chunk_dataset_cpu.py) and, through it, against the reference class.
// The arg-max over the track lengths is an ordered bucket structure (first index among the longest tracks == np.argmax).
#include <cstdint>
#include <cstring>
#include <set>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/dfsfm_b200.h"

namespace dfsfm {
void set_last_error(const std::string& s);

namespace {

class PySet {
  public:
    enum : uint8_t { UNUSED = 0, ACTIVE = 1, DUMMY = 2 };
    struct Entry {
        int64_t key;
        uint8_t st;
    };
    PySet() : table_(8, Entry{0, UNUSED}), mask_(7), fill_(0), used_(0) {}
    static PySet from_list(const std::vector<int64_t>& keys) {  // set(iterable): one set_add_key per element
        PySet s;
        for (int64_t k : keys) s.add(k);
        return s;
    }
    static PySet copy_of(const PySet& o) {  // set(a_set) / set_copy: make_new_set -> set_update_internal -> set_merge
        PySet s;
        s.merge(o);
        return s;
    }
    size_t size() const { return used_; }
    bool contains(int64_t key) const {
        const Entry* e = look(key);
        return e->st == ACTIVE;
    }
    void add(int64_t key) {  // set_add_entry
        check_key(key);
        size_t perturb = static_cast<size_t>(key);
        size_t i = static_cast<size_t>(key) & mask_;
        Entry* freeslot = nullptr;
        while (true) {
            Entry* e = &table_[i];
            int probes = (i + kLinearProbes <= mask_) ? kLinearProbes : 0;
            do {
                if (e->st == UNUSED) goto found_unused_or_dummy_;
                if (e->st == ACTIVE) {
                    if (e->key == key) return;  // found_active
                } else {
                    freeslot = e;
                }
                ++e;
            } while (probes--);
            perturb >>= kPerturbShift;
            i = (i * 5 + 1 + perturb) & mask_;
            continue;
        found_unused_or_dummy_:
            if (freeslot != nullptr) {
                ++used_;
                freeslot->key = key;
                freeslot->st = ACTIVE;
                return;
            }
            ++fill_;
            ++used_;
            e->key = key;
            e->st = ACTIVE;
            if (fill_ * 5 < mask_ * 3) return;
            resize(used_ > 50000 ? used_ * 2 : used_ * 4);
            return;
        }
    }
    void discard(int64_t key) {  // set_discard_entry: the slot becomes a dummy
        Entry* e = const_cast<Entry*>(look(key));
        if (e->st != ACTIVE) return;
        e->st = DUMMY;
        --used_;
    }
    // set_merge(so = *this, other): |= and the copy constructor
    void merge(const PySet& o) {
        if (&o == this || o.used_ == 0) return;
        if ((fill_ + o.used_) * 5 >= mask_ * 3) resize((used_ + o.used_) * 2);
        if (fill_ == 0 && mask_ == o.mask_ && o.fill_ == o.used_) {  // same geometry, no dummies: slots are copied as they are
            for (size_t i = 0; i <= o.mask_; ++i)
                if (o.table_[i].st == ACTIVE) table_[i] = o.table_[i];
            fill_ = o.fill_;
            used_ = o.used_;
            return;
        }
        if (fill_ == 0) {
            fill_ = used_ = o.used_;
            for (size_t i = 0; i <= o.mask_; ++i)
                if (o.table_[i].st == ACTIVE) insert_clean(table_, mask_, o.table_[i].key);
            return;
        }
        for (size_t i = 0; i <= o.mask_; ++i)
            if (o.table_[i].st == ACTIVE) add(o.table_[i].key);
    }
    // set_difference_update_internal: -=
    void difference_update(const PySet& o) {
        if (&o == this) { *this = PySet(); return; }
        for (size_t i = 0; i <= o.mask_; ++i)
            if (o.table_[i].st == ACTIVE) discard(o.table_[i].key);
        if ((fill_ - used_) <= mask_ / 4) return;  // more than a quarter dummies: resize them away
        resize(used_ > 50000 ? used_ * 2 : used_ * 4);
    }
    // set_difference(so = a, other = b): a - b
    static PySet difference(const PySet& a, const PySet& b) {
        if ((a.size() >> 2) > b.size()) {  // set_copy_and_difference
            PySet r = copy_of(a);
            r.difference_update(b);
            return r;
        }
        PySet r;
        for (size_t i = 0; i <= a.mask_; ++i)
            if (a.table_[i].st == ACTIVE && !b.contains(a.table_[i].key)) r.add(a.table_[i].key);
        return r;
    }
    // set_intersection(so = a, other = b): a & b -- iterates the smaller operand (b on ties)
    static PySet intersection(const PySet& a, const PySet& b) {
        if (&a == &b) return copy_of(a);
        const PySet* so = &a;
        const PySet* other = &b;
        if (other->size() > so->size()) std::swap(so, other);
        PySet r;
        for (size_t i = 0; i <= other->mask_; ++i)
            if (other->table_[i].st == ACTIVE && so->contains(other->table_[i].key)) r.add(other->table_[i].key);
        return r;
    }
    std::vector<int64_t> to_list() const {  // list(a_set): slot order
        std::vector<int64_t> out;
        out.reserve(used_);
        for (size_t i = 0; i <= mask_; ++i)
            if (table_[i].st == ACTIVE) out.push_back(table_[i].key);
        return out;
    }

  private:
    static constexpr int kLinearProbes = 9;
    static constexpr int kPerturbShift = 5;
    std::vector<Entry> table_;
    size_t mask_, fill_, used_;

    static void check_key(int64_t key) {
        if (key < 0 || key >= (int64_t(1) << 60)) throw std::runtime_error("bag assignment: image ids must be non-negative (hash(n) == n is assumed)");
    }
    const Entry* look(int64_t key) const {  // set_lookkey
        size_t perturb = static_cast<size_t>(key);
        size_t i = static_cast<size_t>(key) & mask_;
        while (true) {
            const Entry* e = &table_[i];
            int probes = (i + kLinearProbes <= mask_) ? kLinearProbes : 0;
            do {
                if (e->st == UNUSED) return e;
                if (e->st == ACTIVE && e->key == key) return e;
                ++e;
            } while (probes--);
            perturb >>= kPerturbShift;
            i = (i * 5 + 1 + perturb) & mask_;
        }
    }
    static void insert_clean(std::vector<Entry>& table, size_t mask, int64_t key) {  // set_insert_clean
        size_t perturb = static_cast<size_t>(key);
        size_t i = static_cast<size_t>(key) & mask;
        while (true) {
            Entry* e = &table[i];
            if (e->st == UNUSED) { e->key = key; e->st = ACTIVE; return; }
            if (i + kLinearProbes <= mask) {
                for (int j = 0; j < kLinearProbes; ++j) {
                    ++e;
                    if (e->st == UNUSED) { e->key = key; e->st = ACTIVE; return; }
                }
            }
            perturb >>= kPerturbShift;
            i = (i * 5 + 1 + perturb) & mask;
        }
    }
    void resize(size_t minused) {  // set_table_resize: smallest power of two > minused, entries re-inserted in slot order
        size_t newsize = 8;
        while (newsize <= minused) newsize <<= 1;
        std::vector<Entry> nt(newsize, Entry{0, UNUSED});
        for (size_t i = 0; i <= mask_; ++i)
            if (table_[i].st == ACTIVE) insert_clean(nt, newsize - 1, table_[i].key);
        table_.swap(nt);
        mask_ = newsize - 1;
        fill_ = used_;
    }
};

struct Bag {
    std::vector<int64_t> image_ids;
    std::vector<int64_t> track_ids;
    std::vector<int64_t> ref_img;                  // per track
    std::vector<std::vector<int64_t>> query_imgs;  // per track
};

// FeatureTrackStatus (construct_matching_data.py:10-163) over flat inputs.
class TrackStatus {
  public:
    TrackStatus(int64_t n_tracks, const int64_t* track_ids, const int64_t* ref_img_ids, const int64_t* obs_offset, const int64_t* obs_img_ids,
                int max_track_length, int max_num_img_in_bag)
        : n_(n_tracks), ids_(track_ids, track_ids + n_tracks), ref_(ref_img_ids, ref_img_ids + n_tracks), max_len_(max_track_length),
          max_bag_(max_num_img_in_bag) {
        query_.resize(n_);
        len_.resize(n_);
        idx_of_.reserve(static_cast<size_t>(n_) * 2);
        int64_t longest = 1;
        for (int64_t t = 0; t < n_; ++t) {
            idx_of_[ids_[t]] = t;
            // query_img_ids = list(set(image_ids) - {assigned_img_id})                                   (:33-36)
            std::vector<int64_t> obs(obs_img_ids + obs_offset[t], obs_img_ids + obs_offset[t + 1]);
            PySet all = PySet::from_list(obs);
            PySet one;
            one.add(ref_[t]);
            query_[t] = PySet::difference(all, one).to_list();
            len_[t] = static_cast<int64_t>(query_[t].size()) + 1;
            total_ += len_[t] - 1;
            if (len_[t] > longest) longest = len_[t];
        }
        buckets_.resize(static_cast<size_t>(longest) + 1);
        for (int64_t t = 0; t < n_; ++t) buckets_[len_[t]].insert(t);
        cur_max_ = longest;
    }
    bool not_empty() const { return total_ != 0; }  // `is_empty()` of the reference returns total_track_length != 0
    int64_t argmax() {                               // np.argmax(track_length): first index among the longest
        while (cur_max_ > 0 && buckets_[cur_max_].empty()) --cur_max_;
        return *buckets_[cur_max_].begin();
    }
    int64_t id_of(int64_t t) const { return ids_[t]; }
    int64_t ref_of(int64_t t) const { return ref_[t]; }
    bool has(int64_t track_id) const { return idx_of_.count(track_id) != 0; }
    int64_t index(int64_t track_id) const { return idx_of_.at(track_id); }
    int64_t length(int64_t t) const { return len_[t]; }
    const std::vector<int64_t>& query(int64_t t) const { return query_[t]; }
    int max_bag() const { return max_bag_; }
    // get_images_from_trackID_and_update (:61-82): the first max_track_length-1 query images of the longest track
    std::vector<int64_t> take_for_bag(int64_t t) {
        std::vector<int64_t> out;
        const int64_t cap = max_len_ - 1;
        if (static_cast<int64_t>(query_[t].size()) > cap) {
            out.assign(query_[t].begin(), query_[t].begin() + cap);
            query_[t].erase(query_[t].begin(), query_[t].begin() + cap);
            total_ -= cap;
            set_len(t, len_[t] - cap);
        } else {
            out.swap(query_[t]);
            total_ -= len_[t] - 1;
            set_len(t, 1);
        }
        return out;
    }
    // update_track_status (:91-102): query = list(set(query) - set(excluded))
    void exclude(int64_t t, const PySet& excluded) {
        PySet q = PySet::from_list(query_[t]);
        PySet ex = PySet::copy_of(excluded);
        query_[t] = PySet::difference(q, ex).to_list();
        total_ -= static_cast<int64_t>(excluded.size());
        set_len(t, len_[t] - static_cast<int64_t>(excluded.size()));
    }

  private:
    int64_t n_;
    std::vector<int64_t> ids_, ref_;
    int max_len_, max_bag_;
    std::vector<std::vector<int64_t>> query_;
    std::vector<int64_t> len_;
    std::unordered_map<int64_t, int64_t> idx_of_;
    std::vector<std::set<int64_t>> buckets_;
    int64_t cur_max_ = 1;
    int64_t total_ = 0;
    void set_len(int64_t t, int64_t l) {
        buckets_[len_[t]].erase(t);
        len_[t] = l;
        buckets_[l].insert(t);
    }
};

struct FrameDict {  // colmap image id -> track ids whose reference node lies on that image (keyframe_dict)
    std::unordered_map<int64_t, std::pair<int64_t, int64_t>> range;
    const int64_t* tracks = nullptr;
};

// get_relevant_tracks (:104-157)
void relevant_tracks(TrackStatus& st, const FrameDict& fd, int64_t exclude_track_id, Bag& bag) {
    for (size_t bi = 0; bi < bag.image_ids.size(); ++bi) {  // the list grows inside the loop and the loop sees the additions
        const int64_t image_id = bag.image_ids[bi];
        auto it = fd.range.find(image_id);
        if (it == fd.range.end()) throw std::runtime_error("bag assignment: image id without a keyframe_dict entry");
        for (int64_t k = it->second.first; k < it->second.second; ++k) {
            const int64_t track_id = fd.tracks[k];
            if (track_id == exclude_track_id) continue;
            if (!st.has(track_id)) continue;  // processed by other workers
            const int64_t t = st.index(track_id);
            if (st.length(t) == 1) continue;  // already empty
            if (st.ref_of(t) != image_id) throw std::runtime_error("bag assignment: keyframe_dict / assignment mismatch");
            const std::vector<int64_t>& query = st.query(t);
            PySet q1 = PySet::from_list(query), b1 = PySet::from_list(bag.image_ids);
            PySet common = PySet::intersection(q1, b1);
            PySet q2 = PySet::from_list(query), b2 = PySet::from_list(bag.image_ids);
            PySet excl = PySet::difference(q2, b2);
            const int64_t add_quota = st.max_bag() - static_cast<int64_t>(bag.image_ids.size());
            if (add_quota > 0 && excl.size() != 0) {
                std::vector<int64_t> extra = excl.to_list();
                if (static_cast<int64_t>(extra.size()) > add_quota) extra.resize(static_cast<size_t>(add_quota));
                bag.image_ids.insert(bag.image_ids.end(), extra.begin(), extra.end());
                excl.difference_update(PySet::from_list(extra));
                common.merge(PySet::from_list(extra));
            }
            // (len(common) == len(query) or len(exclude) >= 0) and len(common) != 0
            if (common.size() != 0) {
                std::vector<int64_t> common_list = common.to_list();
                st.exclude(t, common);
                bag.track_ids.push_back(track_id);
                bag.ref_img.push_back(image_id);
                bag.query_imgs.push_back(std::move(common_list));
            }
        }
    }
}

struct Result {
    std::vector<int64_t> bag_img_off, bag_img;          // CSR over chunked bags
    std::vector<int64_t> bag_trk_off, trk_id, trk_ref;  // CSR over chunked bags -> tracks
    std::vector<int64_t> trk_q_off, trk_q;              // CSR over tracks -> query image ids
};

}  // namespace
}  // namespace dfsfm

struct dfsfm_bags {
    dfsfm::Result r;
};

extern "C" {

int dfsfm_assign_bags(dfsfm_bags_t** out, int64_t n_tracks, const int64_t* track_ids, const int64_t* ref_img_ids, const int64_t* obs_offset,
                      const int64_t* obs_img_ids, int64_t n_frames, const int64_t* frame_img_ids, const int64_t* frame_offset,
                      const int64_t* frame_track_ids, int max_track_length, int max_num_img_in_bag, int chunk) {
    using namespace dfsfm;
    try {
        if (max_num_img_in_bag <= 0) max_num_img_in_bag = max_track_length;
        TrackStatus st(n_tracks, track_ids, ref_img_ids, obs_offset, obs_img_ids, max_track_length, max_num_img_in_bag);
        FrameDict fd;
        fd.tracks = frame_track_ids;
        for (int64_t f = 0; f < n_frames; ++f) fd.range[frame_img_ids[f]] = {frame_offset[f], frame_offset[f + 1]};
        std::vector<Bag> bags;
        while (st.not_empty()) {  // assign_bags (:226-261)
            const int64_t t = st.argmax();
            Bag bag;
            const int64_t ref = st.ref_of(t);
            std::vector<int64_t> query = st.take_for_bag(t);
            bag.image_ids.push_back(ref);
            bag.image_ids.insert(bag.image_ids.end(), query.begin(), query.end());
            bag.track_ids.push_back(st.id_of(t));
            bag.ref_img.push_back(ref);
            bag.query_imgs.push_back(query);
            relevant_tracks(st, fd, st.id_of(t), bag);
            bags.push_back(std::move(bag));
        }
        auto* h = new dfsfm_bags;
        Result& r = h->r;
        r.bag_img_off.push_back(0);
        r.bag_trk_off.push_back(0);
        r.trk_q_off.push_back(0);
        auto emit = [&](const Bag& b, const std::vector<size_t>& sel) {
            r.bag_img.insert(r.bag_img.end(), b.image_ids.begin(), b.image_ids.end());
            r.bag_img_off.push_back(static_cast<int64_t>(r.bag_img.size()));
            for (size_t i : sel) {
                r.trk_id.push_back(b.track_ids[i]);
                r.trk_ref.push_back(b.ref_img[i]);
                r.trk_q.insert(r.trk_q.end(), b.query_imgs[i].begin(), b.query_imgs[i].end());
                r.trk_q_off.push_back(static_cast<int64_t>(r.trk_q.size()));
            }
            r.bag_trk_off.push_back(static_cast<int64_t>(r.trk_id.size()));
        };
        for (const Bag& b : bags) {  // chunk_bags (:202-224) with chunks_balance (round robin)
            const size_t n = b.track_ids.size();
            if (chunk > 0 && n > static_cast<size_t>(chunk)) {
                const size_t n_split = n / static_cast<size_t>(chunk) + 1;
                for (size_t s = 0; s < n_split; ++s) {
                    std::vector<size_t> sel;
                    for (size_t i = s; i < n; i += n_split) sel.push_back(i);
                    emit(b, sel);
                }
            } else {
                std::vector<size_t> sel(n);
                for (size_t i = 0; i < n; ++i) sel[i] = i;
                emit(b, sel);
            }
        }
        *out = h;
        return 0;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return 1;
    }
}

void dfsfm_bags_sizes(const dfsfm_bags_t* h, int64_t* n_bags, int64_t* n_bag_images, int64_t* n_tracks, int64_t* n_query) {
    *n_bags = static_cast<int64_t>(h->r.bag_img_off.size()) - 1;
    *n_bag_images = static_cast<int64_t>(h->r.bag_img.size());
    *n_tracks = static_cast<int64_t>(h->r.trk_id.size());
    *n_query = static_cast<int64_t>(h->r.trk_q.size());
}

void dfsfm_bags_export(const dfsfm_bags_t* h, int64_t* bag_img_off, int64_t* bag_img, int64_t* bag_trk_off, int64_t* trk_id, int64_t* trk_ref,
                       int64_t* trk_q_off, int64_t* trk_q) {
    const dfsfm::Result& r = h->r;
    auto cp = [](int64_t* dst, const std::vector<int64_t>& v) { if (!v.empty()) memcpy(dst, v.data(), v.size() * sizeof(int64_t)); };
    cp(bag_img_off, r.bag_img_off); cp(bag_img, r.bag_img); cp(bag_trk_off, r.bag_trk_off); cp(trk_id, r.trk_id); cp(trk_ref, r.trk_ref);
    cp(trk_q_off, r.trk_q_off); cp(trk_q, r.trk_q);
}

void dfsfm_bags_destroy(dfsfm_bags_t* h) { delete h; }

// test hook: one CPython-set expression on small integers, result in iteration order.  op: 0 list(set(a)), 1 list(set(a) - set(b)),
// 2 list(set(a) & set(b)), 3 s = set(a); s |= set(b); list(s), 4 s = set(a); s -= set(b); list(s), 5 list(set(set(a)))
int dfsfm_debug_pyset(int op, const int64_t* a, int64_t na, const int64_t* b, int64_t nb, int64_t* out, int64_t* n_out) {
    using namespace dfsfm;
    try {
        PySet sa = PySet::from_list(std::vector<int64_t>(a, a + na)), sb = PySet::from_list(std::vector<int64_t>(b, b + nb));
        PySet r;
        switch (op) {
            case 0: r = sa; break;
            case 1: r = PySet::difference(sa, sb); break;
            case 2: r = PySet::intersection(sa, sb); break;
            case 3: r = sa; r.merge(sb); break;
            case 4: r = sa; r.difference_update(sb); break;
            case 5: r = PySet::copy_of(sa); break;
            default: throw std::runtime_error("bad op");
        }
        const std::vector<int64_t> l = r.to_list();
        memcpy(out, l.data(), l.size() * sizeof(int64_t));
        *n_out = static_cast<int64_t>(l.size());
        return 0;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return 1;
    }
}

}  // extern "C"
