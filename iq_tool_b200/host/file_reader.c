/* file_reader.c — see file_reader.h. */
#include "file_reader.h"

#include <unistd.h>

#include "app_context.h"
#include "constants.h"
#include "log.h"
#include "pipeline_types.h"
#include "queue.h"
#include "ring_buffer.h"
#include "signal_handler.h"

void iqgpu_file_reader_loop(ModuleContext *ctx, SNDFILE *capture, const char *kind)
{
    AppResources *resources = ctx->resources;
    const bool paced = resources->pacing_is_required;
    const size_t high_water = paced ? (size_t)(ring_buffer_get_capacity(resources->writer_input_buffer) * IO_WRITER_BUFFER_HIGH_WATER_MARK) : 0;

    while (!is_shutdown_requested() && !resources->error_occurred) {
        if (paced && ring_buffer_get_size(resources->writer_input_buffer) > high_water) {
            usleep(10000);                                  /* the writer is behind: let it drain */
            continue;
        }
        SampleChunk *chunk = (SampleChunk *)queue_dequeue(resources->free_sample_chunk_queue);
        if (!chunk) break;
        chunk->stream_discontinuity_event = false;

        const int64_t got = sf_read_raw(capture, chunk->raw_input_data, (sf_count_t)chunk->raw_input_capacity_bytes);
        if (got < 0) {
            log_fatal("%s read error on the input file.", kind);
            pthread_mutex_lock(&resources->progress_mutex);
            resources->error_occurred = true;
            pthread_mutex_unlock(&resources->progress_mutex);
            request_shutdown();
            queue_enqueue(resources->free_sample_chunk_queue, chunk);
            break;
        }
        chunk->frames_read = got / (int64_t)resources->input_bytes_per_sample_pair;
        chunk->packet_sample_format = resources->input_format;
        chunk->is_last_chunk = chunk->frames_read == 0;
        if (!chunk->is_last_chunk) {
            pthread_mutex_lock(&resources->progress_mutex);
            resources->total_frames_read += (unsigned long long)chunk->frames_read;
            pthread_mutex_unlock(&resources->progress_mutex);
        }
        if (!queue_enqueue(resources->reader_output_queue, chunk)) {
            queue_enqueue(resources->free_sample_chunk_queue, chunk);
            break;
        }
        if (chunk->is_last_chunk) break;
    }
}
