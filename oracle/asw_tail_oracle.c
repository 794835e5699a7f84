/* placeholder translation unit: the post-hot-path tail (refinement, median) is added
 * in a later step; see oracle/asw_oracle.c for the header that applies to oracle/. */
typedef int asw_tail_oracle_placeholder;
