/* sndfile_min.c — see sndfile_min.h.  sf_read_raw / sf_seek with libsndfile's meaning for a PCM data chunk:
 * raw reads are whole frames and stop at the end of the chunk, seeks are in frames. */
#define _FILE_OFFSET_BITS 64
#include "sndfile_min.h"

#include <stdlib.h>

#include "iqgpu.h"

SNDFILE *sfmin_open(const char *path, uint64_t data_offset, uint64_t data_bytes, uint32_t frame_bytes)
{
    if (!frame_bytes) return NULL;
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    if (fseeko(f, (off_t)data_offset, SEEK_SET) != 0) { fclose(f); return NULL; }
    SNDFILE *s = (SNDFILE *)calloc(1, sizeof(*s));
    if (!s) { fclose(f); return NULL; }
    s->file = f;
    s->data_offset = data_offset;
    s->data_bytes = data_bytes - data_bytes % frame_bytes;
    s->frame_bytes = frame_bytes;
    return s;
}

void sfmin_close(SNDFILE *s)
{
    if (!s) return;
    if (s->file) fclose(s->file);
    free(s);
}

sf_count_t sf_read_raw(SNDFILE *s, void *ptr, sf_count_t bytes)
{
    if (!s || !ptr || bytes < 0) return -1;
    uint64_t want = (uint64_t)bytes;
    const uint64_t left = s->data_bytes - s->position;
    if (want > left) want = left;
    want -= want % s->frame_bytes;
    if (!want) return 0;
    const size_t got = fread(ptr, 1, (size_t)want, s->file);
    if (got < want && ferror(s->file)) return -1;
    const size_t whole = got - got % s->frame_bytes;
    if (whole != got) fseeko(s->file, (off_t)(s->data_offset + s->position + whole), SEEK_SET);
    s->position += whole;
    return (sf_count_t)whole;
}

sf_count_t sf_seek(SNDFILE *s, sf_count_t frames, int whence)
{
    if (!s) return -1;
    const int64_t total = (int64_t)(s->data_bytes / s->frame_bytes), now = (int64_t)(s->position / s->frame_bytes);
    int64_t target = whence == SEEK_SET ? frames : whence == SEEK_CUR ? now + frames : whence == SEEK_END ? total + frames : -1;
    if (target < 0 || target > total) return -1;
    if (fseeko(s->file, (off_t)(s->data_offset + (uint64_t)target * s->frame_bytes), SEEK_SET) != 0) return -1;
    s->position = (uint64_t)target * s->frame_bytes;
    return target;
}

SNDFILE *sfmin_create(const char *path, int container, int sample_format, int sample_rate_hz)
{
    unsigned char header[80];
    const size_t n = iqgpu_wav_header_bytes(container);
    if (!n || iqgpu_wav_build_header(container, sample_format, sample_rate_hz, 0, header, sizeof(header)) != IQGPU_OK) return NULL;
    FILE *f = fopen(path, "wb");
    if (!f) return NULL;
    SNDFILE *s = (SNDFILE *)calloc(1, sizeof(*s));
    if (!s || fwrite(header, 1, n, f) != n) { fclose(f); free(s); return NULL; }
    s->file = f;
    s->writing = 1;
    s->container = container;
    s->sample_format = sample_format;
    s->sample_rate_hz = sample_rate_hz;
    return s;
}

sf_count_t sfmin_write_raw(SNDFILE *s, const void *ptr, sf_count_t bytes)
{
    if (!s || !s->writing || bytes < 0) return 0;
    const size_t done = fwrite(ptr, 1, (size_t)bytes, s->file);
    s->bytes_written += done;
    return (sf_count_t)done;
}

int sfmin_finish(SNDFILE *s)
{
    if (!s) return -1;
    int rc = 0;
    if (s->writing) {
        unsigned char header[80];
        const size_t n = iqgpu_wav_header_bytes(s->container);
        if (iqgpu_wav_build_header(s->container, s->sample_format, s->sample_rate_hz, s->bytes_written, header, sizeof(header)) != IQGPU_OK ||
            fseeko(s->file, 0, SEEK_SET) != 0 || fwrite(header, 1, n, s->file) != n)
            rc = -1;
    }
    if (fclose(s->file) != 0) rc = -1;
    free(s);
    return rc;
}
