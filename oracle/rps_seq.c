/* Sequential rock-paper-scissors pair loop -- CPU oracle, TEST INFRASTRUCTURE ONLY.
 *
 * Restates /root/reference/interaction_simulator.py:104-105 (the in-place loop over pairs)
 * calling /root/reference/interactions.py:13-40 (the rule), with the random draw of pair k
 * supplied as u[k] (consumed only if the two species differ, interactions.py:17-20).
 * Checked against the unmodified Python reference through tests/golden/rps_*.npz.
 */
#include <stdint.h>

#define ROCK 1      /* interactions.py:5 */
#define PAPER 2
#define SCISSORS 3

int64_t rps_sequential(int8_t *species, const int64_t *pairs, const double *u, int64_t P,
                       double pRS, double pPR, double pSP)
{
    int64_t draws = 0;
    for (int64_t k = 0; k < P; ++k) {
        const int64_t p1 = pairs[2 * k], p2 = pairs[2 * k + 1];
        const int8_t s1 = species[p1], s2 = species[p2];
        if (s1 == s2) continue;                         /* :17 */
        const double r = u[k];                          /* :20 */
        ++draws;
        int winner = 0;                                 /* 0 = None, 1 = p1, 2 = p2   (:22) */
        if      (s1 == ROCK     && s2 == SCISSORS) winner = (r < pRS) ? 1 : 2;   /* :24-25 */
        else if (s1 == ROCK     && s2 == PAPER)    winner = (r < pPR) ? 2 : 1;   /* :26-27 */
        else if (s1 == PAPER    && s2 == ROCK)     winner = (r < pPR) ? 1 : 2;   /* :28-29 */
        else if (s1 == PAPER    && s2 == SCISSORS) winner = (r < pSP) ? 2 : 1;   /* :30-31 */
        else if (s1 == SCISSORS && s2 == ROCK)     winner = (r < pRS) ? 2 : 1;   /* :32-33 */
        else if (s1 == SCISSORS && s2 == PAPER)    winner = (r < pSP) ? 1 : 2;   /* :34-35 */
        if (winner == 1) species[p2] = s1;              /* :37-38 */
        else if (winner == 2) species[p1] = s2;         /* :39-40 */
    }
    return draws;
}
