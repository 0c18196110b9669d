// api_more.cuh -- remaining extern "C" entry points (Bpr mini-batch, GeoIE, score+top-K)
#pragma once
extern "C" {
int poi_bpr_train_batch(poi_engine* e, float*, int64_t, float*, int64_t, int32_t, const int32_t*, const int32_t*,
                        const int32_t*, const int32_t*, int64_t, float, float, double*) {
    POI_FAIL(e, "poi_bpr_train_batch: not implemented yet");
}
int poi_geoie_train(poi_engine* e, const poi_geoie_params*, int32_t, const int32_t*, const int32_t*, int32_t,
                    const float*, const float*, const int32_t*, int32_t, float, float, double*) {
    POI_FAIL(e, "poi_geoie_train: not implemented yet");
}
int poi_score_topk(poi_engine* e, const float*, int32_t, const float*, int64_t, int32_t, const float*, float,
                   int32_t, int32_t*) {
    POI_FAIL(e, "poi_score_topk: not implemented yet");
}
}
