// placeholder
int bf16_pack_weights(SrhepHandle* h, const float*) { return fail(h, SRHEP_E_INVALID, "bf16 path not built yet"); }
void bf16_free_weights(SrhepHandle*) {}
int bf16_on_bind(SrhepHandle*) { return 0; }
void bf16_forward(Engine& E, const Pass&, const int*) { E.rc = fail(E.h, SRHEP_E_INVALID, "bf16 path not built yet"); }
int64_t default_pass_tokens(int) { return 65536; }
