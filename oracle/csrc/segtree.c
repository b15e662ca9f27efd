/* ORACLE (test infrastructure): C restatement of the reference's array-heap segment trees.
 *
 * Follows agilerl/components/segment_tree.py:
 *   __setitem__  :81-95   leaf at cap+idx, then parent[i] = op(tree[2i], tree[2i+1]) up to the root
 *   __getitem__  :97-108
 *   operate      :28-79   recursive range query [start, end)
 *   retrieve     :136-156 go left iff tree[left] > upperbound (strict), else subtract and go right
 * and the PER arithmetic of agilerl/components/replay_buffer.py:
 *   _update_priority   :311-329   leaf = priority ** alpha (C pow == Python float pow)
 *   _sample_proportional :357-381 stratified upper bounds in fp64 from f32 uniforms
 *   _calculate_weights :383-409
 *
 * All arithmetic is IEEE double, identical bit-for-bit to CPython's float arithmetic.
 * Built by oracle/build.py with -O2 -ffp-contract=off (no FMA contraction).
 */
#include <math.h>
#include <stdint.h>

/* op: 0 = add (sum tree), 1 = min */
static inline double opf(int op, double a, double b) { return op == 0 ? a + b : (b < a ? b : a); }
/* note: Python's min(a, b) returns a unless b < a  -> same rule above */

void ost_init(double *tree, int64_t cap, int op) {
    double v = op == 0 ? 0.0 : INFINITY;
    for (int64_t i = 0; i < 2 * cap; ++i) tree[i] = v;
}

void ost_set(double *tree, int64_t cap, int op, int64_t idx, double val) {
    idx += cap;
    tree[idx] = val;
    idx /= 2;
    while (idx >= 1) {
        tree[idx] = opf(op, tree[2 * idx], tree[2 * idx + 1]);
        idx /= 2;
    }
}

double ost_get(const double *tree, int64_t cap, int64_t idx) { return tree[cap + idx]; }

static double helper(const double *tree, int op, int64_t start, int64_t end, int64_t node,
                     int64_t ns, int64_t ne) {
    if (start == ns && end == ne) return tree[node];
    int64_t mid = (ns + ne) / 2;
    if (end <= mid) return helper(tree, op, start, end, 2 * node, ns, mid);
    if (mid + 1 <= start) return helper(tree, op, start, end, 2 * node + 1, mid + 1, ne);
    return opf(op, helper(tree, op, start, mid, 2 * node, ns, mid),
               helper(tree, op, mid + 1, end, 2 * node + 1, mid + 1, ne));
}

double ost_operate(const double *tree, int64_t cap, int op, int64_t start, int64_t end) {
    if (end <= 0) end += cap;
    end -= 1;
    return helper(tree, op, start, end, 1, 0, cap - 1);
}

int64_t ost_retrieve(const double *tree, int64_t cap, double ub) {
    int64_t idx = 1;
    while (idx < cap) {
        int64_t left = 2 * idx;
        if (tree[left] > ub) idx = left;
        else { ub -= tree[left]; idx = left + 1; }
    }
    return idx - cap;
}

/* batched PER helpers (sequential semantics, exactly the reference's loops) */
void oper_update(double *sum_tree, double *min_tree, int64_t cap, const int64_t *idx,
                 const double *priority, int64_t n, double alpha, double *max_priority) {
    for (int64_t i = 0; i < n; ++i) {
        double pa = pow(priority[i], alpha);
        ost_set(sum_tree, cap, 0, idx[i], pa);
        ost_set(min_tree, cap, 1, idx[i], pa);
        if (priority[i] > *max_priority) *max_priority = priority[i];
    }
}

void oper_sample(const double *sum_tree, int64_t cap, const float *uniforms, int64_t B,
                 int64_t *out_idx) {
    double total = sum_tree[1];
    double segment = total / (double)B;
    for (int64_t i = 0; i < B; ++i) {
        double a = segment * (double)i;
        double b = segment * (double)(i + 1);
        double ub = (double)uniforms[i] * (b - a) + a;
        out_idx[i] = ost_retrieve(sum_tree, cap, ub);
    }
}

void oper_weights(const double *sum_tree, const double *min_tree, int64_t cap, const int64_t *idx,
                  int64_t B, double beta, int64_t size, float *out_w) {
    double total = sum_tree[1];
    double p_min = min_tree[1] / total;
    double max_weight = pow(p_min * (double)size, -beta);
    for (int64_t i = 0; i < B; ++i) {
        double p = sum_tree[cap + idx[i]] / total;
        double w = pow(p * (double)size, -beta);
        out_w[i] = (float)(w / max_weight);
    }
}
