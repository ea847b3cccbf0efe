/*
 * gm_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's algorithm for the rollout hot path's
 * integer / fp64 part: legacy numpy MT19937 draws, the 3-regular topology
 * generator, integer all-pairs shortest paths, the Routing environment
 * (reset / step / observations) and the SimpleEnvironment.  Each function cites
 * the reference file:line it follows (paths relative to /root/reference).
 *
 * Pinned against outputs of the unmodified reference recorded by
 * tools/gen_golden.py into tests/golden/ (the reference itself has no golden
 * vectors for this path).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* MT19937 exactly as numpy's legacy RandomState drives it   */
/* ------------------------------------------------------------------------ */
typedef struct {
    uint32_t key[624];
    int pos;
} gmo_rng;

/* np.random.seed(int) == init_genrand (numpy/random/src/mt19937/mt19937.c) */
void gmo_rng_seed(gmo_rng *r, uint32_t seed) {
    for (int i = 0; i < 624; i++) {
        r->key[i] = seed;
        seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
    }
    r->pos = 624;
}

static void gmo_rng_gen(gmo_rng *r) {
    uint32_t *mt = r->key;
    for (int k = 0; k < 624; k++) {
        uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    r->pos = 0;
}

uint32_t gmo_rng_u32(gmo_rng *r) {
    if (r->pos >= 624) gmo_rng_gen(r);
    uint32_t y = r->key[r->pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* np.random.random() == genrand_res53 */
double gmo_rng_double(gmo_rng *r) {
    uint32_t a = gmo_rng_u32(r) >> 5, b = gmo_rng_u32(r) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

/* np.random.randint(high) / shuffle's interval: masked rejection on one u32 per
 * try; an interval of width 0 consumes nothing. Returns a value in [0, max]. */
uint32_t gmo_rng_interval(gmo_rng *r, uint32_t max) {
    if (max == 0) return 0;
    uint32_t mask = max;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    if (max == 0xffffffffu) return gmo_rng_u32(r);
    uint32_t v;
    do { v = gmo_rng_u32(r) & mask; } while (v > max);
    return v;
}

uint32_t gmo_rng_randint(gmo_rng *r, uint32_t high) { return gmo_rng_interval(r, high - 1); }

/* np.random.shuffle on a 1-d int array / list */
void gmo_rng_shuffle(gmo_rng *r, int32_t *a, int n) {
    for (int i = n - 1; i >= 1; i--) {
        int j = (int)gmo_rng_interval(r, (uint32_t)i);
        int32_t t = a[i]; a[i] = a[j]; a[j] = t;
    }
}

int gmo_rng_sizeof(void) { return (int)sizeof(gmo_rng); }
int gmo_rng_pos(const gmo_rng *r) { return r->pos; }

/* ------------------------------------------------------------------------ */
/* topology generator -- src/env/network.py:122-213                          */
/* ------------------------------------------------------------------------ */
typedef struct { double d; int j; } gmo_dist;

static int gmo_dist_cmp(const void *a, const void *b) {
    const gmo_dist *x = a, *y = b;
    if (x->d < y->d) return -1;
    if (x->d > y->d) return 1;
    return x->j - y->j; /* list.sort is stable: ties keep ascending j */
}

/* One call of _create_random_topology (network.py:122-195) followed by
 * _check_topology_constraints (:197-213). Returns 1 if valid. Outputs:
 *   edges[3*E] = (start,end,length) in creation order, n_edges,
 *   node_edges[3*N] edge ids sorted by neighbour id (only meaningful if valid),
 *   nbr_creation[3*N] neighbours in creation order (-1 padded), xy[2*N]. */
int gmo_topology_try(gmo_rng *r, int N, int32_t *edges, int32_t *n_edges_out,
                     int32_t *node_edges, int32_t *nbr_creation, double *xy) {
    int *deg = calloc(N, sizeof(int));
    int32_t *ne = malloc(sizeof(int32_t) * 3 * N);
    gmo_dist *dis = malloc(sizeof(gmo_dist) * N);
    for (int i = 0; i < 3 * N; i++) { nbr_creation[i] = -1; ne[i] = -1; }
    for (int i = 0; i < N; i++) { /* :134-138 x then y */
        xy[2 * i] = gmo_rng_double(r);
        xy[2 * i + 1] = gmo_rng_double(r);
    }
    int t_edge = 0;
    for (int i = 0; i < N; i++) {
        for (int j = 0; j < N; j++) { /* :143-150 */
            double dx = xy[2 * j] - xy[2 * i], dy = xy[2 * j + 1] - xy[2 * i + 1];
            dis[j].d = dx * dx + dy * dy;
            dis[j].j = j;
        }
        qsort(dis, N, sizeof(gmo_dist), gmo_dist_cmp); /* :153 */
        for (int j = 1; j < N; j++) {                   /* :157-188 */
            if (deg[i] == 3) break;
            int c = dis[j].j;
            int already = 0;
            for (int q = 0; q < deg[c]; q++) already |= (nbr_creation[3 * c + q] == i);
            if (deg[c] < 3 && !already) {
                int len = (int)(sqrt(dis[j].d) * 10.0) / 2 + 1; /* :173 */
                edges[3 * t_edge] = i < c ? i : c;
                edges[3 * t_edge + 1] = i < c ? c : i;
                edges[3 * t_edge + 2] = len;
                ne[3 * c + deg[c]] = t_edge;
                ne[3 * i + deg[i]] = t_edge;
                nbr_creation[3 * i + deg[i]++] = c;
                nbr_creation[3 * c + deg[c]++] = i;
                t_edge++;
            }
        }
    }
    *n_edges_out = t_edge;
    int valid = 1;
    for (int i = 0; i < N; i++) valid &= (deg[i] == 3); /* :205-207 */
    if (valid) { /* connectivity :210 */
        int *seen = calloc(N, sizeof(int)), *stack = malloc(sizeof(int) * N), sp = 0, cnt = 1;
        seen[0] = 1; stack[sp++] = 0;
        while (sp) {
            int u = stack[--sp];
            for (int q = 0; q < 3; q++) {
                int v = nbr_creation[3 * u + q];
                if (!seen[v]) { seen[v] = 1; cnt++; stack[sp++] = v; }
            }
        }
        valid = (cnt == N);
        free(seen); free(stack);
    }
    if (valid) { /* :191-195 sort each node's edges by the neighbour's id */
        for (int i = 0; i < N; i++) {
            int32_t e[3], o[3];
            for (int q = 0; q < 3; q++) {
                e[q] = ne[3 * i + q];
                o[q] = edges[3 * e[q]] == i ? edges[3 * e[q] + 1] : edges[3 * e[q]];
            }
            for (int a = 0; a < 3; a++)
                for (int b = a + 1; b < 3; b++)
                    if (o[b] < o[a]) { int32_t t = o[a]; o[a] = o[b]; o[b] = t; t = e[a]; e[a] = e[b]; e[b] = t; }
            for (int q = 0; q < 3; q++) node_edges[3 * i + q] = e[q];
        }
    }
    free(deg); free(ne); free(dis);
    return valid;
}

static int gmo_in_set(int64_t v, const int64_t *set, int n) {
    for (int i = 0; i < n; i++) if (set[i] == v) return 1;
    return 0;
}

/* _create_valid_network (network.py:215-272). `global` is the caller's stream.
 * seed_mode: 0 = draw a fresh seed from `global` (no list), 1 = use given_seed.
 * Returns the topology seed finally used; *repetitions as in :244-255. */
int64_t gmo_topology_create(gmo_rng *global, int N, int seed_mode, int64_t given_seed,
                            const int64_t *exclude, int n_exclude, int32_t *edges,
                            int32_t *node_edges, int32_t *nbr_creation, double *xy,
                            int32_t *repetitions) {
    int64_t seed = given_seed;
    if (seed_mode == 0) { /* :229-232 */
        do { seed = gmo_rng_randint(global, 2147483647u); } while (gmo_in_set(seed, exclude, n_exclude));
    }
    gmo_rng local; /* :241-242: state saved, generator reseeded -> use a private stream */
    gmo_rng_seed(&local, (uint32_t)seed);
    *repetitions = 0;
    for (;;) {
        int32_t ne;
        int ok = gmo_topology_try(&local, N, edges, &ne, node_edges, nbr_creation, xy);
        (*repetitions)++;
        if (ok) break;
        if (seed_mode != 0) return -1; /* :251 "Provided seed is invalid" */
        do { seed = gmo_rng_randint(&local, 2147483647u); } while (gmo_in_set(seed, exclude, n_exclude));
        gmo_rng_seed(&local, (uint32_t)seed);
    }
    return seed;
}

/* shortest path WEIGHTS (network.py:274-290): sums of integer edge lengths */
void gmo_apsp(int N, int E, const int32_t *edges, int32_t *apsp) {
    const int32_t INF = 1 << 29;
    for (int i = 0; i < N * N; i++) apsp[i] = INF;
    for (int i = 0; i < N; i++) apsp[i * N + i] = 0;
    for (int e = 0; e < E; e++) {
        int a = edges[3 * e], b = edges[3 * e + 1], w = edges[3 * e + 2];
        if (w < apsp[a * N + b]) { apsp[a * N + b] = w; apsp[b * N + a] = w; }
    }
    for (int k = 0; k < N; k++)
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++)
                if (apsp[i * N + k] + apsp[k * N + j] < apsp[i * N + j])
                    apsp[i * N + j] = apsp[i * N + k] + apsp[k * N + j];
}

/* ------------------------------------------------------------------------ */
/* Routing environment -- src/env/routing.py                                 */
/* ------------------------------------------------------------------------ */
typedef struct {
    int32_t N, A, E, env_var, k, congestion, action_mask, ttl, eval_info, VW;
    /* topology (one graph) */
    const int32_t *node_edges; /* [N,3] */
    const int32_t *edges;      /* [E,3] start,end,length */
    const int32_t *apsp;       /* [N,N] */
    /* packet state, routing.py:12-40 */
    int32_t *now, *target, *edge, *time, *ttl_left, *spw, *start, *agent_steps;
    double *size;
    uint32_t *visited; /* [A,VW] */
    double *load;      /* [E] */
    uint8_t *mask;     /* [A,4] */
    /* eval-info accumulators (routing.py:167-169) */
    double *sum_packets_per_node, *sum_packets_per_edge;
} gmo_env;

static inline int gmo_other(const gmo_env *e, int edge, int node) { /* network.py:31-39 */
    return e->edges[3 * edge] == node ? e->edges[3 * edge + 1] : e->edges[3 * edge];
}

/* routing.py:119-144 with the three draws supplied by the caller */
static void gmo_reset_packet(gmo_env *e, int i, int start, int target, double size) {
    if (e->edge[i] != -1) e->load[e->edge[i]] -= e->size[i]; /* :126-127 */
    e->now[i] = start; e->target[i] = target; e->size[i] = size; e->start[i] = start;
    e->time[i] = 0; e->edge[i] = -1; e->ttl_left[i] = e->ttl;
    e->spw[i] = e->apsp[start * e->N + target];
    for (int w = 0; w < e->VW; w++) e->visited[i * e->VW + w] = 0;
    e->visited[i * e->VW + start / 32] |= 1u << (start % 32);
    if (e->action_mask) { /* :140-144 */
        e->mask[4 * i] = (start != target);
        e->mask[4 * i + 1] = e->mask[4 * i + 2] = e->mask[4 * i + 3] = 0;
    }
}

/* routing.py:160-178 (network.reset() is the caller's business) */
void gmo_routing_reset(gmo_env *e, const int32_t *d_start, const int32_t *d_target,
                       const double *d_size) {
    for (int i = 0; i < e->A; i++) e->agent_steps[i] = 0;
    for (int j = 0; j < e->E; j++) e->load[j] = 0.0;
    if (e->eval_info) {
        for (int j = 0; j < e->N; j++) e->sum_packets_per_node[j] = 0;
        for (int j = 0; j < e->E; j++) e->sum_packets_per_edge[j] = 0;
    }
    for (int i = 0; i < e->A; i++) {
        e->edge[i] = -1; /* fresh Data(i): edge=-1, so nothing is freed */
        gmo_reset_packet(e, i, d_start[i], d_target[i], d_size[i]);
    }
}

/* routing.py:360-495. Draw slot s is consumed by the s-th reset in id order.
 * Outputs per agent: reward f32, done, delays (agent_steps at done else 0),
 * arrived, spr (f64), looped; info = {looped, throughput, dropped, blocked};
 * extra (eval info) = {total_edge_load, occupied_edges, packets_on_edges,
 * total_packet_size}. Returns the number of draw slots consumed. */
int gmo_routing_step(gmo_env *e, const int32_t *act, const int32_t *d_start,
                     const int32_t *d_target, const double *d_size, float *reward,
                     uint8_t *done, int32_t *delays, uint8_t *arrived, double *spr,
                     uint8_t *looped, int32_t *info, double *extra, int32_t *packet_dist) {
    int A = e->A, blocked = 0, slot = 0;
    for (int i = 0; i < A; i++) {
        reward[i] = 0.0f; done[i] = 0; delays[i] = 0; arrived[i] = 0; spr[i] = 0.0; looped[i] = 0;
        e->agent_steps[i] += 1; /* :371 */
    }
    for (int i = 0; i < A; i++) { /* :380-412 */
        if (e->eval_info && e->edge[i] == -1) e->sum_packets_per_node[e->now[i]] += 1;
        if (e->edge[i] == -1 && act[i] != 0) {
            int t = e->node_edges[3 * e->now[i] + act[i] - 1];
            if (e->congestion && e->load[t] + e->size[i] > 1) {
                reward[i] -= 0.2f; /* float32 op, :398 */
                blocked++;
            } else {
                e->edge[i] = t;
                e->time[i] = e->edges[3 * t + 2];
                e->load[t] += e->size[i];
                int nn = gmo_other(e, t, e->now[i]);
                e->now[i] = nn;
                uint32_t *v = &e->visited[i * e->VW + nn / 32], bit = 1u << (nn % 32);
                if (*v & bit) looped[i] = 1; else *v |= bit;
            }
        }
    }
    if (e->eval_info) { /* :414-441 */
        double tl = 0, tps = 0; int occ = 0, poe = 0;
        for (int j = 0; j < e->E; j++) { tl += e->load[j]; occ += e->load[j] > 0; }
        for (int i = 0; i < A; i++) {
            if (e->edge[i] != -1) { e->sum_packets_per_edge[e->edge[i]] += 1; poe++; }
            tps += e->size[i];
            packet_dist[i] = e->apsp[e->now[i] * e->N + e->target[i]];
        }
        extra[0] = tl; extra[1] = occ; extra[2] = poe; extra[3] = tps;
    }
    int n_looped = 0, n_success = 0, n_dropped = 0;
    for (int i = 0; i < A; i++) { /* :444-491 */
        e->ttl_left[i] -= 1;
        if (e->edge[i] != -1) {
            e->time[i] -= 1;
            if (e->time[i] <= 0) { e->load[e->edge[i]] -= e->size[i]; e->edge[i] = -1; }
        }
        int drop = (e->ttl > 0 && e->ttl_left[i] <= 0);
        if (e->action_mask) { /* :456-469 */
            uint8_t *m = &e->mask[4 * i];
            if (e->edge[i] != -1) { m[0] = m[1] = m[2] = m[3] = 0; }
            else {
                m[0] = 1;
                for (int q = 0; q < 3; q++) {
                    int o = gmo_other(e, e->node_edges[3 * e->now[i] + q], e->now[i]);
                    m[1 + q] = (e->visited[i * e->VW + o / 32] >> (o % 32)) & 1u;
                }
                if (m[0] + m[1] + m[2] + m[3] == 4) drop = 1;
            }
        }
        int reached = (e->edge[i] == -1 && e->now[i] == e->target[i]);
        if (reached || drop) {
            reward[i] += reached ? 10.0f : -10.0f; /* :474 */
            done[i] = 1;
            int opt = e->spw[i] > 1 ? e->spw[i] : 1; /* :479 */
            if (reached) { arrived[i] = 1; spr[i] = (double)e->agent_steps[i] / (double)opt; n_success++; }
            else n_dropped++;
            delays[i] = e->agent_steps[i];
            e->agent_steps[i] = 0;
            gmo_reset_packet(e, i, d_start[slot], d_target[slot], d_size[slot]);
            slot++;
        }
        n_looped += looped[i];
    }
    info[0] = n_looped; info[1] = n_success; info[2] = n_dropped; info[3] = blocked;
    return slot;
}

/* routing.py:187-235 -> out [N, 4N+8] f32 */
void gmo_routing_node_obs(const gmo_env *e, float *out) {
    int N = e->N, W = 4 * N + 8;
    memset(out, 0, sizeof(float) * (size_t)N * W);
    for (int j = 0; j < N; j++) {
        float *o = out + (size_t)j * W;
        o[j] = 1.0f;
        int np_ = 0; double tl = 0;
        for (int i = 0; i < e->A; i++)
            if (e->now[i] == j && e->edge[i] == -1) { np_++; tl += e->size[i]; }
        o[N] = (float)np_; o[N + 1] = (float)tl;
        for (int q = 0; q < 3; q++) {
            int k = e->node_edges[3 * j + q];
            float *s = o + N + 2 + q * (N + 2);
            s[gmo_other(e, k, j)] = 1.0f;
            s[N] = (float)e->edges[3 * k + 2];
            s[N + 1] = (float)e->load[k];
        }
    }
}

int gmo_routing_obs_width(const gmo_env *e) {
    int N = e->N, W = 6 * N + 10;
    if (e->env_var == 2) W += 5 * e->k;
    if (e->env_var == 3) W += N * N + N * (4 * N + 8);
    return W;
}

static int gmo_is_nbr_or_same(const gmo_env *e, int a, int b) {
    if (a == b) return 1;
    for (int q = 0; q < 3; q++)
        if (gmo_other(e, e->node_edges[3 * a + q], a) == b) return 1;
    return 0;
}

/* routing.py:269-358 -> out [A, W] f32 */
void gmo_routing_obs(const gmo_env *e, float *out) {
    int N = e->N, A = e->A, W = gmo_routing_obs_width(e);
    memset(out, 0, sizeof(float) * (size_t)A * W);
    float *glob = NULL;
    if (e->env_var == 3) {
        glob = malloc(sizeof(float) * (size_t)N * (4 * N + 8));
        gmo_routing_node_obs(e, glob);
    }
    for (int i = 0; i < A; i++) {
        float *o = out + (size_t)i * W;
        int now = e->now[i];
        o[now] = 1.0f;
        o[N + e->target[i]] = 1.0f;
        o[2 * N] = (float)(e->edge[i] != -1);
        if (e->edge[i] != -1) o[2 * N + 1 + gmo_other(e, e->edge[i], now)] = 1.0f;
        o[3 * N + 1] = (float)e->time[i];
        o[3 * N + 2] = (float)e->size[i];
        o[3 * N + 3] = (float)i;
        for (int q = 0; q < 3; q++) {
            int k = e->node_edges[3 * now + q];
            float *s = o + 3 * N + 4 + q * (N + 2);
            s[gmo_other(e, k, now)] = 1.0f;
            s[N] = (float)e->edges[3 * k + 2];
            s[N + 1] = (float)e->load[k];
        }
        int p = 6 * N + 10;
        if (e->env_var == 2) { /* :328-343 */
            int count = 0;
            for (int j = 0; j < A && count < e->k; j++) {
                if (j == i || !gmo_is_nbr_or_same(e, now, e->now[j])) continue;
                o[p++] = (float)e->now[j]; o[p++] = (float)e->target[j];
                o[p++] = (float)e->edge[j]; o[p++] = (float)e->size[j]; o[p++] = (float)i;
                count++;
            }
            for (; count < e->k; count++) for (int q = 0; q < 5; q++) o[p++] = -1.0f;
        }
        if (e->env_var == 3) { /* :271-275, 353-354 */
            for (int a = 0; a < N; a++)
                for (int b = 0; b < N; b++) o[p++] = (float)gmo_is_nbr_or_same(e, a, b);
            memcpy(o + p, glob, sizeof(float) * (size_t)N * (4 * N + 8));
        }
    }
    free(glob);
}

/* routing.py:522-539 (via the neigh lists built in :318-326) -> [A,A] i8 */
void gmo_routing_adj(const gmo_env *e, int8_t *adj) {
    for (int i = 0; i < e->A; i++)
        for (int j = 0; j < e->A; j++)
            adj[i * e->A + j] = (int8_t)(i == j || gmo_is_nbr_or_same(e, e->now[i], e->now[j]));
}

/* routing.py:256-267 -> [N,A] i8 */
void gmo_routing_node_agent(const gmo_env *e, int8_t *m) {
    memset(m, 0, (size_t)e->N * e->A);
    for (int a = 0; a < e->A; a++) m[e->now[a] * e->A + a] = 1;
}

/* network.py:385-389 -> [N,N] i8 */
void gmo_node_adj(int N, int E, const int32_t *edges, int8_t *adj) {
    memset(adj, 0, (size_t)N * N);
    for (int i = 0; i < N; i++) adj[i * N + i] = 1;
    for (int k = 0; k < E; k++) {
        adj[edges[3 * k] * N + edges[3 * k + 1]] = 1;
        adj[edges[3 * k + 1] * N + edges[3 * k]] = 1;
    }
}

/* ------------------------------------------------------------------------ */
/* SimpleEnvironment -- src/env/simple_environment.py:106-187, 289-315       */
/* ------------------------------------------------------------------------ */
/* _build_network with the global stream. Outputs: scores[3], edges[2][2],
 * start_node, start_edge_order[2] (router[n0].edge). */
void gmo_simple_build(gmo_rng *r, int random_topology, int32_t *scores, int32_t *edges,
                      int32_t *start_node, int32_t *start_edges) {
    int32_t border[2] = {-1, 1};
    gmo_rng_shuffle(r, border, 2); /* :113-115 */
    scores[0] = border[0]; scores[1] = 0; scores[2] = border[1];
    if (random_topology) gmo_rng_shuffle(r, scores, 3); /* :118-120 */
    int n0 = scores[0] == 0 ? 0 : (scores[1] == 0 ? 1 : 2);
    int n1 = (n0 + 1) % 3, n2 = (n1 + 1) % 3;
    *start_node = n0;
    for (int i = 0; i < 3; i++) { gmo_rng_double(r); gmo_rng_double(r); } /* :128-132 positions */
    int32_t dest[2] = {n1, n2};
    if (random_topology) gmo_rng_shuffle(r, dest, 2); /* :139-141 */
    for (int k = 0; k < 2; k++) {                     /* :143-171 */
        int32_t en[2] = {n0, dest[k]};
        if (random_topology) gmo_rng_shuffle(r, en, 2);
        edges[2 * k] = en[0]; edges[2 * k + 1] = en[1];
    }
    start_edges[0] = 0; start_edges[1] = 1;
    if (random_topology && dest[1] < dest[0]) { start_edges[0] = 1; start_edges[1] = 0; } /* argsort :174-179 */
}

/* step (:289-315): reward = score of the chosen neighbour of the start node */
int gmo_simple_step(const int32_t *scores, const int32_t *edges, int start_node,
                    const int32_t *start_edges, int act) {
    int t = start_edges[act];
    int dst = edges[2 * t] == start_node ? edges[2 * t + 1] : edges[2 * t];
    return scores[dst];
}

/* ------------------------------------------------------------------------ */
/* batched drivers (used by the CPU baseline leg of bench.py)                */
/* ------------------------------------------------------------------------ */
/* Advances envs [lo, hi) of B independent envs that share one topology; arrays
 * are [B, ...] contiguous.  Threading is the caller's (oracle.py runs disjoint
 * [lo, hi) ranges on a thread pool; ctypes releases the GIL). */
typedef struct {
    gmo_env proto; /* pointers = base of the [B,...] arrays */
    int32_t B;
} gmo_batch;

static void gmo_env_at(const gmo_batch *b, int i, gmo_env *e) {
    *e = b->proto;
    int A = e->A, E = e->E, VW = e->VW;
    e->now += (size_t)i * A; e->target += (size_t)i * A; e->edge += (size_t)i * A;
    e->time += (size_t)i * A; e->ttl_left += (size_t)i * A; e->spw += (size_t)i * A;
    e->start += (size_t)i * A; e->agent_steps += (size_t)i * A; e->size += (size_t)i * A;
    e->visited += (size_t)i * A * VW; e->load += (size_t)i * E; e->mask += (size_t)i * A * 4;
    if (e->sum_packets_per_node) { e->sum_packets_per_node += (size_t)i * e->N; e->sum_packets_per_edge += (size_t)i * E; }
}

void gmo_batch_reset(const gmo_batch *b, int lo, int hi, const int32_t *ds, const int32_t *dt, const double *dz) {
    for (int i = lo; i < hi && i < b->B; i++) {
        gmo_env e; gmo_env_at(b, i, &e);
        size_t o = (size_t)i * e.A;
        gmo_routing_reset(&e, ds + o, dt + o, dz + o);
    }
}

void gmo_batch_step(const gmo_batch *b, int lo, int hi, const int32_t *act, const int32_t *ds, const int32_t *dt,
                    const double *dz, float *reward, uint8_t *done, int32_t *delays,
                    uint8_t *arrived, double *spr, uint8_t *looped, int32_t *info,
                    int32_t *n_resets) {
    for (int i = lo; i < hi && i < b->B; i++) {
        gmo_env e; gmo_env_at(b, i, &e);
        size_t o = (size_t)i * e.A;
        double extra[4]; int32_t pd[4096];
        n_resets[i] = gmo_routing_step(&e, act + o, ds + o, dt + o, dz + o, reward + o, done + o,
                                       delays + o, arrived + o, spr + o, looped + o, info + 4 * i,
                                       extra, pd);
    }
}

void gmo_batch_observe(const gmo_batch *b, int lo, int hi, float *obs, int8_t *adj, float *node_obs, int8_t *node_agent) {
    for (int i = lo; i < hi && i < b->B; i++) {
        gmo_env e; gmo_env_at(b, i, &e);
        int W = gmo_routing_obs_width(&e);
        if (obs) gmo_routing_obs(&e, obs + (size_t)i * e.A * W);
        if (adj) gmo_routing_adj(&e, adj + (size_t)i * e.A * e.A);
        if (node_obs) gmo_routing_node_obs(&e, node_obs + (size_t)i * e.N * (4 * e.N + 8));
        if (node_agent) gmo_routing_node_agent(&e, node_agent + (size_t)i * e.N * e.A);
    }
}
